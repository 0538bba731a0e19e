"""Marching-cubes spec sanity on the CPU oracle (own spec): the generated 256-case table yields a
closed, consistently oriented surface whose normals point to the observed (sdf>0) side."""
import ctypes as C

import numpy as np

import oracle_bindings as ob
from scannet_b200 import synth
from scannet_b200._lib import TsdfParams


def test_table_properties(built):
    L = ob.tsdf_oracle_lib()
    n = np.zeros(256, np.uint8); t = np.zeros((256, 36), np.uint8)
    L.oracle_mc_table(n.ctypes.data, t.ctypes.data)
    assert n[0] == 0 and n[255] == 0 and n.max() <= 12
    def ends(e):
        a, j = e >> 2, e & 3; u, v = j & 1, j >> 1
        off = [(0, u, v), (u, 0, v), (u, v, 0)][a]
        c0 = off[0] | off[1] << 1 | off[2] << 2
        return c0, c0 | (1 << a)
    for cs in range(256):
        used = set(t[cs, : 3 * n[cs]].tolist())
        cut = {e for e in range(12) if ((cs >> ends(e)[0]) & 1) != ((cs >> ends(e)[1]) & 1)}
        assert used == cut, cs                         # every sign-changing edge, and only those
        # each directed edge of the triangulation interior appears with its reverse at most once -> loops are simple
        de = {}
        for k in range(n[cs]):
            tri = t[cs, 3 * k: 3 * k + 3]
            for i in range(3):
                key = (int(tri[i]), int(tri[(i + 1) % 3])); de[key] = de.get(key, 0) + 1
        assert all(v == 1 for v in de.values()), cs
    # single inside corner: one triangle whose normal points away from that corner
    for c in range(8):
        cs = 1 << c
        assert n[cs] == 1
        pts = []
        for e in t[cs, :3]:
            c0, c1 = ends(int(e)); p0 = np.array([c0 & 1, c0 >> 1 & 1, c0 >> 2], float); p1 = np.array([c1 & 1, c1 >> 1 & 1, c1 >> 2], float)
            pts.append((p0 + p1) / 2)
        nrm = np.cross(pts[1] - pts[0], pts[2] - pts[0])
        corner = np.array([c & 1, c >> 1 & 1, c >> 2], float)
        assert nrm @ (np.mean(pts, 0) - corner) > 0


def test_oracle_mesh_is_closed_and_oriented(built):
    p = TsdfParams(); p.voxel_size = 0.004; p.trunc_base = 0.02; p.trunc_scale = 0.01; p.depth_min = 0.1; p.depth_max = 6.0
    p.max_integration_distance = 4.0; p.weight_sample = 1; p.weight_max = 255; p.width = 96; p.height = 72; p.depth_shift = 1000.0
    D, Cc, P, K = synth.make_frames(3, seed=2, width=96, height=72, loop_frames=300)
    o = ob.OracleTsdf(p, threads=4)
    for i in range(3):
        o.integrate(D[i], Cc[i], P[i], K)
    xyz, rgb, tri = o.extract_mesh()
    assert len(tri) > 1000 and tri.max() == len(xyz) - 1 and len(np.unique(tri)) == len(xyz)
    # consistent orientation: no directed edge twice; interior edges have their reverse
    de = (tri[:, [0, 1, 2]].astype(np.int64) << 32 | tri[:, [1, 2, 0]]).ravel()
    assert len(np.unique(de)) == len(de)
    rev = (tri[:, [1, 2, 0]].astype(np.int64) << 32 | tri[:, [0, 1, 2]]).ravel()
    interior = np.isin(de, rev).mean()
    assert interior > 0.9                                    # the rest is the open boundary of the observed region
    # normals face the camera (sdf>0 side)
    cam = P[1][:3, 3]
    cen = xyz[tri].mean(1); nrm = np.cross(xyz[tri[:, 1]] - xyz[tri[:, 0]], xyz[tri[:, 2]] - xyz[tri[:, 0]])
    facing = ((cam - cen) * nrm).sum(1) > 0
    assert facing.mean() > 0.9
