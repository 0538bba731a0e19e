"""Host-side mirror of the reference Segmentator seams
(/root/reference/Segmentator/segmentator.cpp:71 ``segment_graph``, :123 ``segment``, :253 ``writeToJSON``)
over the C ABI.  All compute happens in libscannet_b200.so (CUDA); nothing here falls back to the CPU."""
from __future__ import annotations

import ctypes as C

import numpy as np

from ._lib import check, lib

EDGE_DTYPE = np.dtype([("w", "<f4"), ("a", "<i4"), ("b", "<i4")])
DEVICE_UNIONFIND = 1          # SCN_SEG_DEVICE_UNIONFIND: Kruskal pass on the GPU (speculative-window kernel) instead of the host loop


def test_depth_flag(d: int) -> int:
    return (d & 0xFF) << 8


def segment_mesh(xyz: np.ndarray, tri: np.ndarray, k_thresh: float = 0.01, seg_min_verts: int = 20, flags: int = 0) -> np.ndarray:
    """segIndices[v] = root vertex id of v's segment (same ids as the reference)."""
    xyz = np.ascontiguousarray(xyz, np.float32); tri = np.ascontiguousarray(tri, np.uint32)
    out = np.zeros(len(xyz), np.int32)
    check(lib().scn_segment_mesh(xyz.ctypes.data_as(C.c_void_p), C.c_uint64(len(xyz)), tri.ctypes.data_as(C.c_void_p),
                                 C.c_uint64(len(tri)), C.c_float(k_thresh), C.c_int32(seg_min_verts),
                                 out.ctypes.data_as(C.c_void_p), C.c_int(flags)))
    return out


def segment_mesh_debug(xyz, tri, k_thresh=0.01, seg_min_verts=20, flags=0):
    """Returns dict(seg, edges_presort, edges_sorted, normals, roots_after_kruskal)."""
    xyz = np.ascontiguousarray(xyz, np.float32); tri = np.ascontiguousarray(tri, np.uint32)
    nV, nF = len(xyz), len(tri)
    out = np.zeros(nV, np.int32); pre = np.zeros(3 * nF, EDGE_DTYPE); srt = np.zeros(3 * nF, EDGE_DTYPE)
    nrm = np.zeros((nV, 3), np.float32); roots = np.zeros(nV, np.int32)
    check(lib().scn_segment_mesh_debug(xyz.ctypes.data_as(C.c_void_p), C.c_uint64(nV), tri.ctypes.data_as(C.c_void_p),
                                       C.c_uint64(nF), C.c_float(k_thresh), C.c_int32(seg_min_verts),
                                       out.ctypes.data_as(C.c_void_p), C.c_int(flags), pre.ctypes.data_as(C.c_void_p),
                                       srt.ctypes.data_as(C.c_void_p), nrm.ctypes.data_as(C.c_void_p),
                                       roots.ctypes.data_as(C.c_void_p)))
    return dict(seg=out, edges_presort=pre, edges_sorted=srt, normals=nrm, roots_after_kruskal=roots)


def segment_graph(n_verts: int, edges: np.ndarray, c: float, flags: int = 0):
    """Sorts `edges` (EDGE_DTYPE) like the reference's std::sort and runs the thresholded Kruskal pass.
    Returns (sorted_edges, roots, sizes)."""
    e = np.ascontiguousarray(edges, EDGE_DTYPE).copy()
    roots = np.zeros(n_verts, np.int32); sizes = np.zeros(n_verts, np.int32)
    check(lib().scn_segment_graph(C.c_int32(n_verts), C.c_int64(len(e)), e.ctypes.data_as(C.c_void_p), C.c_float(c),
                                  roots.ctypes.data_as(C.c_void_p), sizes.ctypes.data_as(C.c_void_p), C.c_int(flags)))
    return e, roots, sizes


def last_timings():
    """(ms[8], sort_kernel_launches): H2D, normals, weights, sort, kruskal, small-merge, gather+D2H+labels, total."""
    ms = (C.c_float * 8)()
    n = check(lib().scn_segment_last_timings(ms))
    return list(ms), n


def last_uf_rounds() -> int:
    """rounds of the device union-find replay in the last segment_mesh call of this thread (0 = host loop)"""
    f = lib().scn_segment_last_uf_rounds
    f.restype = C.c_uint64
    return int(f())
