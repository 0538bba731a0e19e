"""GPU parity: segs.json consumers (scannet_b200/csrc/segs.cu through the C ABI) vs oracle/segs_oracle.c.

Bars: segment ids, vertex lists, object ids: exact.  mLib vertex normals: bit-identical floats.  Surface area: relative 1e-5
(the reference itself sums float areas in unordered_map order — Segmentation.h:121-143 — and the per-face area goes through
acosf/sinf, which differ by ulps between libm and CUDA).  Propagated labels: exact, except vertices the oracle flags as within
1e-5 rad of the normal threshold (acosf again)."""
import numpy as np
import pytest

import oracle_bindings as ob
from scannet_b200 import segmentator, segs, synth

pytestmark = pytest.mark.gpu


def same_floats(a, b):
    a = np.ascontiguousarray(a, np.float32); b = np.ascontiguousarray(b, np.float32)
    return bool(((a.view(np.uint32) == b.view(np.uint32)) | (np.isnan(a) & np.isnan(b))).all())


def check_aggregate(seg, xyz, tri):
    g = segs.aggregate(seg, xyz, tri); o = ob.oracle_segs_aggregate(seg, xyz, tri)
    assert (g["seg_ids"] == o["seg_ids"]).all() and (g["offsets"] == o["offsets"]).all() and (g["vert_ids"] == o["vert_ids"]).all()
    both_nan = np.isnan(g["area"]) & np.isnan(o["area"])
    assert (both_nan | (np.abs(g["area"] - o["area"]) <= 1e-5 * np.abs(o["area"]) + 1e-12)).all()
    return g


def test_aggregate_on_segmented_mesh(built):
    xyz, tri = synth.make_feature_mesh(250, 200, 0)
    seg = segmentator.segment_mesh(xyz, tri)
    g = check_aggregate(seg, xyz, tri)
    assert len(g["seg_ids"]) == len(np.unique(seg)) and g["area"].min() >= 0
    assert int(g["offsets"][-1]) == len(xyz)
    # ids only (no mesh)
    h = segs.aggregate(seg)
    assert h["area"] is None and (h["vert_ids"] == g["vert_ids"]).all()


@pytest.mark.parametrize("case", ["one_segment", "all_distinct", "big_ids", "tiny", "degenerate"])
def test_aggregate_edge_cases(built, case):
    rng = np.random.default_rng(5)
    xyz, tri = synth.make_grid_mesh(40, 30, 1)
    n = len(xyz)
    if case == "one_segment":
        seg = np.zeros(n, np.uint32)
    elif case == "all_distinct":
        seg = rng.permutation(n).astype(np.uint32)
    elif case == "big_ids":
        seg = rng.choice(np.array([0, 1, 0x7FFFFFFF, 0xFFFFFFFF, 0x80000000, 65536], np.uint32), n)
    elif case == "tiny":
        xyz, tri = xyz[:3], np.array([[0, 1, 2]], np.uint32); seg = np.array([4, 4, 4], np.uint32)
    else:
        xyz, tri = synth.make_adversarial_mesh(0); seg = (np.arange(len(xyz)) // 50).astype(np.uint32)
    check_aggregate(seg, xyz, tri)


def test_aggregate_two_million_vertices(built):
    xyz, tri = synth.make_feature_mesh(1600, 1250, 0)
    seg = segmentator.segment_mesh(xyz, tri)
    check_aggregate(seg, xyz, tri)


def test_objects_per_vertex(built):
    rng = np.random.default_rng(2)
    seg = rng.integers(0, 200, 20000).astype(np.uint32)
    groups = [[1, 2, 3], [], [3, 50, 199], [7], [1000], [2]]        # 3 and 2 are claimed twice: the later group wins
    g = segs.objects_per_vertex(seg, groups)
    assert (g == ob.oracle_objects_per_vertex(seg, groups)).all()
    assert set(np.unique(g[seg == 3])) == {3} and set(np.unique(g[seg == 2])) == {6} and (g[seg == 8] == 0).all()
    assert (segs.objects_per_vertex(seg, []) == 0).all()


@pytest.mark.parametrize("mesh", ["feature", "adversarial", "shuffled", "fan"])
def test_vertex_normals_bit_exact(built, mesh):
    if mesh == "feature":
        xyz, tri = synth.make_feature_mesh(250, 200, 3)
    elif mesh == "adversarial":
        xyz, tri = synth.make_adversarial_mesh(0)
    elif mesh == "shuffled":
        xyz, tri = synth.make_grid_mesh(120, 90, 4); tri = tri[np.random.default_rng(1).permutation(len(tri))]
    else:                                                           # one vertex shared by 3000 faces: the heap-sort path
        rng = np.random.default_rng(7); n = 3000
        xyz = np.concatenate([[[0, 0, 0]], rng.normal(size=(n + 1, 3))]).astype(np.float32)
        tri = np.stack([np.zeros(n, np.uint32), np.arange(1, n + 1, dtype=np.uint32), np.arange(2, n + 2, dtype=np.uint32)], 1)
        tri = tri[rng.permutation(n)]
    assert same_floats(segs.vertex_normals(xyz, tri), ob.oracle_vertex_normals_mlib(xyz, tri))


def test_vertex_normals_rejects_bad_index(built):
    from scannet_b200._lib import ScnError
    xyz, tri = synth.make_grid_mesh(10, 10, 0); tri = tri.copy(); tri[3, 1] = len(xyz)
    with pytest.raises(ScnError):
        segs.vertex_normals(xyz, tri)


def _propagation_case(seed, ns=(60, 50), nd=(150, 120), frac_labelled=0.8):
    """decimated mesh = coarse grid, hi-res mesh = fine grid over the same surface, labels = segment-derived objects"""
    sx, st = synth.make_feature_mesh(ns[0], ns[1], seed, spacing=0.05)
    dx, dt = synth.make_feature_mesh(nd[0], nd[1], seed, spacing=0.05 * (ns[0] - 1) / (nd[0] - 1))
    seg = segmentator.segment_mesh(sx, st)
    ids = np.unique(seg); rng = np.random.default_rng(seed)
    groups = [[int(s)] for s in ids if rng.random() < frac_labelled]
    obj = segs.objects_per_vertex(seg, groups)
    return sx, segs.vertex_normals(sx, st), obj, dx, segs.vertex_normals(dx, dt)


@pytest.mark.parametrize("seed,thresh", [(0, 0.5), (1, 0.05), (2, 3.2)])
def test_propagate_labels_matches_oracle(built, seed, thresh):
    sx, sn, so, dx, dn = _propagation_case(seed)
    g = segs.propagate_labels(sx, sn, so, dx, dn, thresh)
    o, edge = ob.oracle_propagate_labels(sx, sn, so, dx, dn, thresh)
    assert edge.mean() < 0.01
    assert (g[~edge] == o[~edge]).all()
    assert (g > 0).mean() > 0.3                                      # the case is not vacuous


def test_propagate_labels_edge_cases(built):
    sx, sn, so, dx, dn = _propagation_case(3, ns=(30, 20), nd=(50, 40))
    # nothing labelled -> all zero; a single labelled vertex; destination far outside the source bbox; NaN destination
    assert (segs.propagate_labels(sx, sn, np.zeros_like(so), dx, dn) == 0).all()
    one = np.zeros_like(so); one[17] = 9
    g = segs.propagate_labels(sx, sn, one, dx, dn, 3.2); o, e = ob.oracle_propagate_labels(sx, sn, one, dx, dn, 3.2)
    assert (g[~e] == o[~e]).all()
    far = dx + np.float32(100.0)
    assert (segs.propagate_labels(sx, sn, so, far, dn) == 0).all()
    bad = dx.copy(); bad[::7, 1] = np.nan
    g = segs.propagate_labels(sx, sn, so, bad, dn); o, e = ob.oracle_propagate_labels(sx, sn, so, bad, dn)
    assert (g[~e] == o[~e]).all() and (g[::7] == 0).all()
    # duplicated source points: equal distances are ordered by source index
    sx2 = np.concatenate([sx, sx]); sn2 = np.concatenate([sn, sn]); so2 = np.concatenate([so, so + 1000 * (so > 0)])
    g = segs.propagate_labels(sx2, sn2, so2, dx, dn, 0.05); o, e = ob.oracle_propagate_labels(sx2, sn2, so2, dx, dn, 0.05)
    assert (g[~e] == o[~e]).all()
