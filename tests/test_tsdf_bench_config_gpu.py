"""GPU parity at exactly the instantiation bench.py times:
640x480, 4 mm voxels, K = 32 frames per block residency (bench.py --batch default; K = 16 in the colour case), depth-only (constant sample weight, stats on), >= 4 batches in
flight so the alloc(k+1) || integrate(k) double buffering is exercised — through BOTH scn_tsdf_integrate_device (frames
resident in HBM) and scn_tsdf_integrate_batch (pinned host frames), plus a colour K=16 case and a reset-and-refuse case.
Tolerance ZERO against oracle/tsdf_oracle.c (own spec, parity unpinned — the reference has no TSDF source)."""
import os

import numpy as np
import pytest

import oracle_bindings as ob
from scannet_b200 import synth, tsdf

pytestmark = pytest.mark.gpu

W, H = 640, 480


def frames(n, seed, loop=1000, color=False):
    D, C, P, K = synth.make_frames(n, seed=seed, width=W, height=H, loop_frames=loop)
    return D, (C if color else None), P, K


def oracle_run(p, D, C, P, K):
    o = ob.OracleTsdf(p, threads=min(os.cpu_count() or 1, 32))
    for i in range(len(D)):
        o.integrate(D[i], None if C is None else C[i], P[i], K)
    ox, ov = o.export()
    oc = o.counters()
    o.close()
    return ox, ov, oc


def check(vol, ox, ov, oc):
    vol.sync()
    gx, gv = vol.download_blocks()
    st = vol.stats()
    assert gx.shape == ox.shape and (gx == ox).all(), "block set differs"
    assert (gv["sdf"].view(np.uint32) == ov["sdf"].view(np.uint32)).all(), f"max |dsdf| = {np.abs(gv['sdf'] - ov['sdf']).max()}"
    assert gv.tobytes() == ov.tobytes()
    assert st.voxels_updated == oc["total_updated"] and st.blocks_visited == oc["total_touched"]
    assert st.frames_integrated == oc["frames_done"]


@pytest.mark.timeout(600)
def test_bench_instantiation_device_and_pinned_batch(built):
    import torch
    n = 144                                                    # 4.5 batches of 32
    D, _, P, K = frames(n, seed=0)
    p = tsdf.default_params(batch_frames=32, max_blocks=1 << 18, hash_slots=1 << 20)     # bench.py's parameters and flags (0: stats on)
    ox, ov, oc = oracle_run(p, D, None, P, K)
    # (1) frames resident in HBM -> scn_tsdf_integrate_device, launched on torch's stream like bench.py
    d = torch.from_numpy(D.view(np.int16)).cuda()
    vol = tsdf.TsdfVolume(p, device=0, stream=torch.cuda.current_stream().cuda_stream)
    vol.integrate_device(n, d.data_ptr(), None, P, K)
    check(vol, ox, ov, oc)
    # (2) same handle after reset, pinned host frames -> scn_tsdf_integrate_batch (H2D double buffer)
    vol.reset()
    h = torch.from_numpy(D.view(np.int16)).pin_memory()
    vol.integrate_batch_ptr(n, h.data_ptr(), None, P, K)
    check(vol, ox, ov, oc)
    # (3) reset again: the proportional clear must leave no residue (a used volume refused from empty gives the same bits)
    vol.reset()
    vol.integrate_device(n, d.data_ptr(), None, P, K)
    check(vol, ox, ov, oc)
    vol.close()


@pytest.mark.timeout(600)
def test_bench_instantiation_split_calls_and_ragged_tail(built):
    """the same stream handed over in uneven calls (37 + 1 + 26 frames): partial batches between full ones"""
    import torch
    n = 64
    D, _, P, K = frames(n, seed=3)
    p = tsdf.default_params(batch_frames=32, max_blocks=1 << 18, hash_slots=1 << 20)
    ox, ov, oc = oracle_run(p, D, None, P, K)
    d = torch.from_numpy(D.view(np.int16)).cuda()
    vol = tsdf.TsdfVolume(p, device=0, stream=torch.cuda.current_stream().cuda_stream)
    o = 0
    for cnt in (37, 1, 26):                                   # 32 + 5 | 1 | 26: partial batches between calls
        vol.integrate_device(cnt, d.data_ptr() + o * W * H * 2, None, P[o:o + cnt], K)
        o += cnt
    check(vol, ox, ov, oc)
    vol.close()


@pytest.mark.timeout(600)
def test_bench_instantiation_colour_k16(built):
    import torch
    n = 48
    D, C, P, K = frames(n, seed=2, color=True)
    p = tsdf.default_params(batch_frames=16, max_blocks=1 << 18, hash_slots=1 << 20)
    ox, ov, oc = oracle_run(p, D, C, P, K)
    d = torch.from_numpy(D.view(np.int16)).cuda(); c = torch.from_numpy(C).cuda()
    vol = tsdf.TsdfVolume(p, device=0, stream=torch.cuda.current_stream().cuda_stream)
    vol.integrate_device(n, d.data_ptr(), c.data_ptr(), P, K)
    check(vol, ox, ov, oc)
    vol.reset()
    vol.integrate_batch(D, C, P, K)
    check(vol, ox, ov, oc)
    vol.close()


@pytest.mark.timeout(600)
def test_bench_instantiation_no_stats_flag(built):
    """SCN_TSDF_NO_STATS compiles a different kernel instantiation: same voxels"""
    import torch
    n = 64
    D, _, P, K = frames(n, seed=5)
    p = tsdf.default_params(batch_frames=32, max_blocks=1 << 18, hash_slots=1 << 20, flags=tsdf.NO_STATS)
    ox, ov, _ = oracle_run(p, D, None, P, K)
    d = torch.from_numpy(D.view(np.int16)).cuda()
    vol = tsdf.TsdfVolume(p, device=0, stream=torch.cuda.current_stream().cuda_stream)
    vol.integrate_device(n, d.data_ptr(), None, P, K)
    vol.sync()
    gx, gv = vol.download_blocks()
    assert (gx == ox).all() and gv.tobytes() == ov.tobytes()
    vol.close()
