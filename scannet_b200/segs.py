"""Host-side mirror of the segs.json consumers of the annotation tools
(/root/reference/AnnotationTools/common/Segmentation.h:57-147, ProjectAnnotations/Visualizer.cpp:259-377,
external/mLib/include/core-mesh/meshData.h:758-782) over the C ABI.  All compute happens in libscannet_b200.so (CUDA)."""
from __future__ import annotations

import ctypes as C

import numpy as np

from ._lib import check, lib


def _take(ptr, n, dtype):
    """copy n items out of a malloc'ed C array and release it"""
    if not ptr:
        return np.zeros(0, dtype)
    ct = {np.uint32: C.c_uint32, np.uint64: C.c_uint64, np.float32: C.c_float}[dtype]
    a = np.ctypeslib.as_array(C.cast(ptr, C.POINTER(ct)), (max(int(n), 1),))[: int(n)].copy()
    lib().scn_free(C.cast(ptr, C.c_void_p))
    return a


def load(path: str):
    """Segmentation::loadFromFile — returns dict(segIndices uint32[n], kThresh, segMinVerts, sceneId)."""
    seg = C.POINTER(C.c_uint32)(); n = C.c_uint64(); k = C.c_float(); m = C.c_uint32(); sid = C.create_string_buffer(512)
    check(lib().scn_segs_load(path.encode(), C.byref(seg), C.byref(n), C.byref(k), C.byref(m), sid, C.c_size_t(512)))
    return dict(segIndices=_take(seg, n.value, np.uint32), kThresh=k.value, segMinVerts=m.value, sceneId=sid.value.decode())


def aggregate(seg: np.ndarray, xyz: np.ndarray | None = None, tri: np.ndarray | None = None):
    """Returns dict(seg_ids uint32[nS] ascending, offsets uint64[nS+1], vert_ids uint32[nV], area float32[nS] or None)."""
    seg = np.ascontiguousarray(seg, np.uint32)
    want = xyz is not None and tri is not None
    if want:
        xyz = np.ascontiguousarray(xyz, np.float32); tri = np.ascontiguousarray(tri, np.uint32)
    ids = C.POINTER(C.c_uint32)(); nS = C.c_uint64(); off = C.POINTER(C.c_uint64)(); vid = C.POINTER(C.c_uint32)(); area = C.POINTER(C.c_float)()
    check(lib().scn_segs_aggregate(seg.ctypes.data_as(C.c_void_p), C.c_uint64(len(seg)),
                                   xyz.ctypes.data_as(C.c_void_p) if want else None, tri.ctypes.data_as(C.c_void_p) if want else None,
                                   C.c_uint64(len(tri) if want else 0), C.byref(ids), C.byref(nS), C.byref(off), C.byref(vid),
                                   C.byref(area) if want else None))
    n = nS.value
    return dict(seg_ids=_take(ids, n, np.uint32), offsets=_take(off, n + 1, np.uint64), vert_ids=_take(vid, len(seg), np.uint32),
                area=_take(area, n, np.float32) if want else None)


def objects_per_vertex(seg: np.ndarray, groups: list) -> np.ndarray:
    """groups[g] = iterable of segment ids (aggregation.json segGroups[g].segments); object id = g + 1, 0 = none."""
    seg = np.ascontiguousarray(seg, np.uint32)
    flat = np.ascontiguousarray(np.concatenate([np.asarray(list(g), np.uint32) for g in groups]) if groups else np.zeros(0, np.uint32), np.uint32)
    offs = np.zeros(len(groups) + 1, np.uint64)
    offs[1:] = np.cumsum([len(list(g)) for g in groups]) if groups else []
    out = np.zeros(len(seg), np.uint32)
    check(lib().scn_segs_objects_per_vertex(seg.ctypes.data_as(C.c_void_p), C.c_uint64(len(seg)), flat.ctypes.data_as(C.c_void_p),
                                            offs.ctypes.data_as(C.c_void_p), C.c_uint64(len(groups)), out.ctypes.data_as(C.c_void_p)))
    return out


def vertex_normals(xyz: np.ndarray, tri: np.ndarray) -> np.ndarray:
    """MeshData::computeVertexNormals, bit-identical to the sequential loop."""
    xyz = np.ascontiguousarray(xyz, np.float32); tri = np.ascontiguousarray(tri, np.uint32)
    out = np.zeros((len(xyz), 3), np.float32)
    check(lib().scn_mesh_vertex_normals(xyz.ctypes.data_as(C.c_void_p), C.c_uint64(len(xyz)), tri.ctypes.data_as(C.c_void_p),
                                        C.c_uint64(len(tri)), out.ctypes.data_as(C.c_void_p)))
    return out


def propagate_labels(src_xyz, src_normals, src_obj, dst_xyz, dst_normals, normal_thresh: float = 0.5) -> np.ndarray:
    """Visualizer::propagateAnnotations with an exact 3-NN search; normal_thresh in radians (zParametersScan.txt:15)."""
    sx = np.ascontiguousarray(src_xyz, np.float32); sn = np.ascontiguousarray(src_normals, np.float32); so = np.ascontiguousarray(src_obj, np.uint32)
    dx = np.ascontiguousarray(dst_xyz, np.float32); dn = np.ascontiguousarray(dst_normals, np.float32)
    out = np.zeros(len(dx), np.uint32)
    check(lib().scn_propagate_labels(sx.ctypes.data_as(C.c_void_p), sn.ctypes.data_as(C.c_void_p), so.ctypes.data_as(C.c_void_p),
                                     C.c_uint64(len(sx)), dx.ctypes.data_as(C.c_void_p), dn.ctypes.data_as(C.c_void_p), C.c_uint64(len(dx)),
                                     C.c_float(normal_thresh), out.ctypes.data_as(C.c_void_p)))
    return out
