"""Host-side mirror of ``ml::SensorData`` (/root/reference/SensReader/c++/src/sensorData.h) over the C ABI:
open / info / per-frame pose, depth (uint16) and colour (RGB8) decode, writer, saveToImages."""
from __future__ import annotations

import ctypes as C

import numpy as np

from ._lib import SensInfo, check, lib


def _L():
    L = lib()
    L.scn_sens_describe.restype = C.c_int64
    return L


class SensFile:
    def __init__(self, path: str | None = None, _handle=None):
        self._h = C.c_void_p()
        if _handle is not None:
            self._h = _handle
        else:
            check(_L().scn_sens_open(str(path).encode(), C.byref(self._h)))
        self.info = SensInfo()
        check(_L().scn_sens_info(self._h, C.byref(self.info)))

    @classmethod
    def create(cls, color_wh, depth_wh, K_color, K_depth, color_compression=0, depth_compression=1, depth_shift=1000.0,
               sensor_name="scannet_b200"):
        h = C.c_void_p()
        kc = np.ascontiguousarray(K_color, np.float32); kd = np.ascontiguousarray(K_depth, np.float32)
        check(_L().scn_sens_create(C.c_uint32(color_wh[0]), C.c_uint32(color_wh[1]), C.c_uint32(depth_wh[0]), C.c_uint32(depth_wh[1]),
                                   kc.ctypes.data_as(C.c_void_p), kd.ctypes.data_as(C.c_void_p), C.c_int32(color_compression),
                                   C.c_int32(depth_compression), C.c_float(depth_shift), sensor_name.encode(), C.byref(h)))
        return cls(_handle=h)

    def close(self):
        if self._h and _L is not None:                    # (module globals are gone when the interpreter is shutting down)
            _L().scn_sens_close(self._h); self._h = C.c_void_p()

    __del__ = close

    def refresh(self):
        check(_L().scn_sens_info(self._h, C.byref(self.info)))

    @property
    def n_frames(self) -> int:
        self.refresh(); return int(self.info.n_frames)

    def K_depth(self) -> np.ndarray:
        return np.array(self.info.depth_intrinsic, np.float32).reshape(4, 4)

    def K_color(self) -> np.ndarray:
        return np.array(self.info.color_intrinsic, np.float32).reshape(4, 4)

    def frame_meta(self, i: int):
        T = np.zeros(16, np.float32); tc = C.c_uint64(); td = C.c_uint64(); cb = C.c_uint64(); db = C.c_uint64()
        check(_L().scn_sens_frame_meta(self._h, C.c_uint64(i), T.ctypes.data_as(C.c_void_p), C.byref(tc), C.byref(td), C.byref(cb), C.byref(db)))
        return T.reshape(4, 4), tc.value, td.value, cb.value, db.value

    def pose(self, i: int) -> np.ndarray:
        return self.frame_meta(i)[0]

    def depth(self, i: int) -> np.ndarray:
        out = np.zeros((self.info.depth_height, self.info.depth_width), np.uint16)
        check(_L().scn_sens_frame_depth_u16(self._h, C.c_uint64(i), out.ctypes.data_as(C.c_void_p)))
        return out

    def decode_depth_device(self, first: int, n: int, d_out_ptr: int, stream: int = 0):
        """Frames [first, first+n) inflated on the GPU straight into device memory at `d_out_ptr` (n*H*W uint16)."""
        check(_L().scn_sens_decode_depth_device(self._h, C.c_uint64(first), C.c_uint32(n), C.c_void_p(d_out_ptr), C.c_void_p(stream)))

    def decode_color_device(self, first: int, n: int, d_out_ptr: int, d_lut_ptr: int = 0, out_px: int = 0, stream: int = 0) -> int:
        """Colour frames [first, first+n) decoded on the GPU into device memory at `d_out_ptr` (RGB8; whole frames, or `out_px`
        pixels per frame picked through the device int32 map `d_lut_ptr`).  Returns the number of frames the device decoded
        (the rest went through the host decoder)."""
        k = C.c_uint32()
        check(_L().scn_sens_decode_color_device(self._h, C.c_uint64(first), C.c_uint32(n), C.c_void_p(d_lut_ptr) if d_lut_ptr else None,
                                                C.c_uint32(out_px), C.c_void_p(d_out_ptr), C.c_void_p(stream), C.byref(k)))
        return k.value

    def read_ahead(self, cache_size: int = 16, n_threads: int = 0):
        """Iterator over (depth uint16 [H,W], colour uint8 [H,W,3], ts_depth, ts_color) decoded by background threads, in stream
        order — the RGBDFrameCacheRead pattern (sensorData.h:1717-1835)."""
        h = C.c_void_p()
        check(_L().scn_sens_cache_create(self._h, C.c_uint32(cache_size), C.c_int(n_threads), C.byref(h)))
        try:
            while True:
                d = np.zeros((self.info.depth_height, self.info.depth_width), np.uint16)
                c = np.zeros((self.info.color_height, self.info.color_width, 3), np.uint8)
                td = C.c_uint64(); tc = C.c_uint64()
                rc = check(_L().scn_sens_cache_next(h, d.ctypes.data_as(C.c_void_p), c.ctypes.data_as(C.c_void_p), C.byref(td), C.byref(tc)))
                if rc == 0:
                    return
                yield d, c, td.value, tc.value
        finally:
            _L().scn_sens_cache_destroy(h)

    def color(self, i: int) -> np.ndarray:
        out = np.zeros((self.info.color_height, self.info.color_width, 3), np.uint8)
        check(_L().scn_sens_frame_color_rgb8(self._h, C.c_uint64(i), out.ctypes.data_as(C.c_void_p)))
        return out

    def set_pose(self, i: int, T: np.ndarray):
        T = np.ascontiguousarray(T, np.float32)
        check(_L().scn_sens_set_pose(self._h, C.c_uint64(i), T.ctypes.data_as(C.c_void_p)))

    def add_frame(self, color, depth: np.ndarray, cam2world: np.ndarray, ts_color=0, ts_depth=0):
        depth = np.ascontiguousarray(depth, np.uint16); T = np.ascontiguousarray(cam2world, np.float32)
        if color is None:
            cp, cn = None, 0
        elif isinstance(color, (bytes, bytearray)):
            buf = (C.c_uint8 * len(color)).from_buffer_copy(color); cp, cn = C.cast(buf, C.c_void_p), len(color)
        else:
            color = np.ascontiguousarray(color, np.uint8); cp, cn = color.ctypes.data_as(C.c_void_p), color.nbytes
        check(_L().scn_sens_add_frame(self._h, cp, C.c_uint64(cn), depth.ctypes.data_as(C.c_void_p), T.ctypes.data_as(C.c_void_p),
                                      C.c_uint64(ts_color), C.c_uint64(ts_depth)))

    def save(self, path: str):
        check(_L().scn_sens_save(self._h, str(path).encode()))

    def save_to_images(self, out_dir: str):
        check(_L().scn_sens_save_to_images(self._h, str(out_dir).encode()))

    def describe(self) -> str:
        buf = C.create_string_buffer(4096)
        check(_L().scn_sens_describe(self._h, buf, C.c_uint64(4096)))
        return buf.value.decode()


def inflate_batch_device(streams, frame_bytes: int, d_out_ptr: int, stream: int = 0):
    """zlib streams (bytes objects) -> frames of `frame_bytes` at device pointer `d_out_ptr`, one warp per stream."""
    n = len(streams)
    bufs = [np.frombuffer(b, np.uint8) if len(b) else np.zeros(1, np.uint8) for b in streams]
    ptrs = (C.c_void_p * max(n, 1))(*[b.ctypes.data for b in bufs])
    lens = (C.c_uint64 * max(n, 1))(*[len(b) for b in streams])
    check(_L().scn_inflate_batch_device(ptrs, lens, C.c_uint32(n), C.c_uint64(frame_bytes), C.c_void_p(d_out_ptr), C.c_void_p(stream)))


def jpeg_decode_batch_device(jpegs, width: int, height: int, d_out_ptr: int, d_lut_ptr: int = 0, out_px: int = 0, stream: int = 0) -> int:
    """JPEG payloads (bytes objects) -> RGB8 frames at device pointer `d_out_ptr`; returns how many the GPU decoded itself."""
    n = len(jpegs)
    bufs = [np.frombuffer(b, np.uint8) if len(b) else np.zeros(1, np.uint8) for b in jpegs]
    ptrs = (C.c_void_p * max(n, 1))(*[b.ctypes.data for b in bufs])
    lens = (C.c_uint64 * max(n, 1))(*[len(b) for b in jpegs])
    k = C.c_uint32()
    check(_L().scn_jpeg_decode_batch_device(ptrs, lens, C.c_uint32(n), C.c_uint32(width), C.c_uint32(height), C.c_void_p(d_lut_ptr) if d_lut_ptr else None,
                                            C.c_uint32(out_px), C.c_void_p(d_out_ptr), C.c_void_p(stream), C.byref(k)))
    return k.value


def inflate_host(data: bytes, cap: int) -> bytes:
    """Host build of the GPU decoder source (test hook)."""
    out = np.zeros(max(cap, 1), np.uint8); n = C.c_size_t()
    check(_L().scn_inflate_host(data, C.c_size_t(len(data)), out.ctypes.data_as(C.c_void_p), C.c_size_t(cap), C.byref(n)))
    return out[: n.value].tobytes()


def inflate_last_timings():
    """(pack_s, kernel_ms, ring_window, n_streams) of this thread's last inflate_batch_device / decode_depth_device"""
    a = C.c_double(); b = C.c_double(); r = C.c_int(); n = C.c_uint32()
    _L().scn_inflate_last_timings(C.byref(a), C.byref(b), C.byref(r), C.byref(n))
    return a.value, b.value, bool(r.value), n.value


def jpeg_last_timings():
    """(host_s, entropy_idct_kernel_ms, colour_kernel_ms) of this thread's last jpeg_decode_batch_device / decode_color_device"""
    a = C.c_double(); b = C.c_double(); c = C.c_double()
    _L().scn_jpeg_last_timings(C.byref(a), C.byref(b), C.byref(c))
    return a.value, b.value, c.value
