"""GPU parity: CUDA Segmentator (through the C ABI) vs the reference.

Bar: BIT-EXACT.  segIndices must equal the ids produced by the unmodified reference compiled with
libstdc++ (committed golden vectors from scripts/make_golden.py, and the C restatement in
oracle/seg_oracle.c which is itself pinned against the reference in test_oracle_pinning.py).
Intermediates are checked too: vertex normals and edge weights as float bit patterns, and the
order of the sorted edge array (libstdc++ introsort's unstable tie order)."""
import os

import numpy as np
import pytest

import oracle_bindings as ob
from scannet_b200 import segmentator, synth

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(__file__), "golden")


def golden():
    return np.load(os.path.join(G, "segmentator_golden.npz"))


def same_floats(a, b):
    """bit-identical, except that any NaN equals any NaN (x86 SSE and CUDA emit different NaN payloads;
    NaNs only arise from zero-area faces, where the reference divides 0 by 0, segmentator.cpp:109-110)."""
    a = np.ascontiguousarray(a); b = np.ascontiguousarray(b)
    return bool(((a.view(np.uint32) == b.view(np.uint32)) | (np.isnan(a) & np.isnan(b))).all())


def same_edges(a, b):
    return same_floats(a["w"], b["w"]) and (a["a"] == b["a"]).all() and (a["b"] == b["b"]).all()


def check_mesh(xyz, tri, k=0.01, m=20):
    dbg = segmentator.segment_mesh_debug(xyz, tri, k, m)
    seg, pre, srt, roots, nrm = ob.oracle_segment(xyz, tri, k, m, want_debug=True)
    assert same_floats(dbg["normals"], nrm), "vertex normals differ"
    assert same_edges(dbg["edges_presort"], pre), "edge weights differ"
    assert same_edges(dbg["edges_sorted"], srt), "sorted edge order differs (introsort tie order)"
    assert (dbg["roots_after_kruskal"] == roots).all()
    assert (dbg["seg"] == seg).all()
    return dbg["seg"]


def test_gates381_golden(built):
    xyz, tri = synth.read_ply(os.path.join(G, "gates381.ply"))
    g = golden()
    for k, m in [(0.01, 20), (0.001, 20), (0.0001, 20), (0.05, 5), (0.01, 1), (0.5, 100)]:
        seg = segmentator.segment_mesh(xyz, tri, k, m)
        assert (seg == g[f"gates381_k{k}_m{m}"]).all(), (k, m)
    check_mesh(xyz, tri)
    assert len(set(segmentator.segment_mesh(xyz, tri).tolist())) == 78


@pytest.mark.parametrize("name,gen", [("grid60x50_s2", lambda: synth.make_grid_mesh(60, 50, 2)),
                                      ("adv_s3", lambda: synth.make_adversarial_mesh(3)),
                                      ("grid250x200_s1", lambda: synth.make_grid_mesh(250, 200, 1))])
def test_synthetic_golden(built, name, gen):
    xyz, tri = gen()
    seg = check_mesh(xyz, tri)
    assert (seg == golden()[name]).all()


def test_feature_mesh_c1(built):
    """C1-sized (50k verts) mesh with steps/bumps: many segments, exact ties, shuffled faces."""
    xyz, tri = synth.make_feature_mesh(250, 200, seed=11)
    seg = check_mesh(xyz, tri)
    assert len(set(seg.tolist())) > 10


def test_edge_cases(built):
    # empty mesh, single triangle, vertices without faces
    assert len(segmentator.segment_mesh(np.zeros((0, 3), np.float32), np.zeros((0, 3), np.uint32))) == 0
    xyz = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0], [5, 5, 5]], np.float32)
    tri = np.array([[0, 1, 2]], np.uint32)
    assert (segmentator.segment_mesh(xyz, tri) == ob.oracle_segment(xyz, tri)).all()
    # high-valence fan (> 32 incident faces per vertex: the heap-sorted corner list path)
    n = 200
    ang = np.linspace(0, 2 * np.pi, n, endpoint=False)
    rng = np.random.default_rng(0)
    ring = np.stack([np.cos(ang), np.sin(ang), 0.05 * rng.normal(size=n)], 1)
    xyz = np.concatenate([[[0, 0, 0.3]], ring]).astype(np.float32)
    tri = np.stack([np.zeros(n, np.int64), 1 + np.arange(n), 1 + (np.arange(n) + 1) % n], 1).astype(np.uint32)
    tri = tri[rng.permutation(n)]
    check_mesh(xyz, tri, 0.05, 3)


def test_out_of_range_index_is_an_error(built):
    from scannet_b200 import ScnError
    xyz = np.zeros((3, 3), np.float32); tri = np.array([[0, 1, 7]], np.uint32)
    with pytest.raises(ScnError):
        segmentator.segment_mesh(xyz, tri)


@pytest.mark.parametrize("n,kind", [(17, "ties"), (300, "ties"), (5000, "rand"), (200000, "ties"), (300000, "few"),
                                    (65536, "organ"), (1 << 20, "rand"), (700001, "ties")])
def test_sort_matches_libstdcxx(built, n, kind):
    """scn_segment_graph's sort must reproduce libstdc++ std::sort's permutation (incl. ties)."""
    rng = np.random.default_rng(n)
    e = np.zeros(n, ob.EDGE_DTYPE)
    if kind == "ties":
        e["w"] = (rng.integers(0, 50, n) / 7.0).astype(np.float32)
    elif kind == "rand":
        e["w"] = rng.random(n).astype(np.float32)
    elif kind == "few":
        e["w"] = rng.integers(0, 3, n).astype(np.float32)
    else:
        e["w"] = np.concatenate([np.arange(n // 2), np.arange(n - n // 2)[::-1]]).astype(np.float32)
    nv = 1000
    e["a"] = rng.integers(0, nv, n); e["b"] = rng.integers(0, nv, n)
    got, roots, sizes = segmentator.segment_graph(nv, e, 0.5)
    want = e.copy()
    ob.seg_oracle().oracle_seg_sort_edges(want.ctypes.data, n)
    assert got.tobytes() == want.tobytes()


@pytest.mark.parametrize("depth", [1, 2, 5, 9])
def test_sort_depth_limit_fallback(built, depth):
    """Force introsort's depth limit low so both tiers hit the heap-sort fallback (stl_algo.h:1924-1928)."""
    n = 300000
    rng = np.random.default_rng(depth)
    e = np.zeros(n, ob.EDGE_DTYPE)
    e["w"] = (rng.integers(0, 1000, n) / 3.0).astype(np.float32); e["a"] = rng.integers(0, 100, n); e["b"] = rng.integers(0, 100, n)
    got, _, _ = segmentator.segment_graph(100, e, 0.5, flags=segmentator.test_depth_flag(depth))
    want = e.copy()
    lib = ob.seg_oracle(); lib.oracle_seg_sort_edges_depth.argtypes = [__import__("ctypes").c_void_p, __import__("ctypes").c_int64, __import__("ctypes").c_int]
    lib.oracle_seg_sort_edges_depth(want.ctypes.data, n, depth)
    assert got.tobytes() == want.tobytes()


def test_segment_graph_golden(built):
    """Reference segment_graph (std::sort + thresholded Kruskal) on a tie-heavy edge list (committed golden)."""
    g = golden()
    got, roots, sizes = segmentator.segment_graph(5000, g["graph_edges_in"], 0.3)
    assert got.tobytes() == g["graph_edges_sorted"].tobytes()
    assert (roots == g["graph_roots"]).all() and (sizes == g["graph_sizes"]).all()


def test_c5_two_million_vertices(built):
    """BASELINE.json configs[4]: 2M-vertex mesh.  Size-independent checks + full oracle comparison."""
    xyz, tri = synth.make_feature_mesh(1600, 1250, seed=5)
    seg = segmentator.segment_mesh(xyz, tri)
    ms, launches = segmentator.last_timings()
    print("C5 timings ms [h2d,normals,weights,sort,kruskal,small,gather+labels,total]:", [round(x, 2) for x in ms], launches)
    # ids are root vertex ids: idempotent, and every root labels itself
    assert (seg[seg] == seg).all()
    ref = ob.oracle_segment(xyz, tri)
    assert (seg == ref).all()


@pytest.mark.parametrize("mesh", ["gates", "grid", "feature", "adversarial"])
def test_device_unionfind_kernel_gives_the_same_ids(built, mesh):
    """SCN_SEG_DEVICE_UNIONFIND: the speculative-window Kruskal kernel (csrc/seg.cu:k_kruskal_window) must leave exactly the
    forest of the sequential loop (segmentator.cpp:71-91): same roots after the Kruskal pass, same final ids."""
    import os
    from scannet_b200 import segmentator, synth
    if mesh == "gates":
        import ctypes as C
        from scannet_b200._lib import check, lib
        path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "gates381.ply").encode()
        px = C.POINTER(C.c_float)(); pt = C.POINTER(C.c_uint32)(); nv = C.c_uint64(); nf = C.c_uint64()
        check(lib().scn_mesh_load(path, C.byref(px), C.byref(nv), C.byref(pt), C.byref(nf)))
        xyz = np.ctypeslib.as_array(px, (nv.value * 3,)).copy().reshape(-1, 3); tri = np.ctypeslib.as_array(pt, (nf.value * 3,)).copy().reshape(-1, 3)
        lib().scn_free(px); lib().scn_free(pt)
    elif mesh == "grid":
        xyz, tri = synth.make_grid_mesh(250, 200, seed=3)
    elif mesh == "feature":
        xyz, tri = synth.make_feature_mesh(300, 260, seed=4)
    else:
        xyz, tri = synth.make_adversarial_mesh(seed=1)
    for kthr in (0.01, 0.0005):
        h = segmentator.segment_mesh_debug(xyz, tri, kthr, 20)
        d = segmentator.segment_mesh_debug(xyz, tri, kthr, 20, flags=segmentator.DEVICE_UNIONFIND)
        assert segmentator.last_uf_rounds() > 0
        assert (h["roots_after_kruskal"] == d["roots_after_kruskal"]).all()
        assert (h["seg"] == d["seg"]).all()
        ref = ob.oracle_segment(xyz, tri, kthr, 20)
        assert (d["seg"] == ref).all()
