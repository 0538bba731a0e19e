"""GPU parity: CUDA TSDF fusion (through the C ABI) vs the scalar CPU statement of TSDF spec v1.

The reference ships no TSDF code — parity here is against this repo's own
oracle (oracle/tsdf_oracle.c, "parity unpinned").  Tolerance: ZERO.  The spec is written in
correctly-rounded binary32 operations + explicit fma, so the block set, every sdf bit pattern,
every colour byte and every weight must be identical."""
import numpy as np
import pytest

import oracle_bindings as ob
from scannet_b200 import synth, tsdf

pytestmark = pytest.mark.gpu


def run_both(p, D, C, P, K, batch_api=True, threads=8):
    vol = tsdf.TsdfVolume(p, device=0)
    if batch_api:
        vol.integrate_batch(D, C, P, K)
    else:
        for i in range(len(D)):
            vol.integrate(D[i], None if C is None else C[i], P[i], K)
    vol.sync()
    gx, gv = vol.download_blocks()
    st = vol.stats()
    vol.close()
    o = ob.OracleTsdf(p, threads=threads)
    for i in range(len(D)):
        o.integrate(D[i], None if C is None else C[i], P[i], K)
    ox, ov = o.export()
    return (gx, gv, st), (ox, ov, o.counters())


def assert_identical(g, o):
    (gx, gv, st), (ox, ov, oc) = g, o
    assert gx.shape == ox.shape, (gx.shape, ox.shape)
    assert (gx == ox).all()
    assert (gv["w"] == ov["w"]).all()
    assert (gv["sdf"].view(np.uint32) == ov["sdf"].view(np.uint32)).all(), \
        f"max |dsdf| = {np.abs(gv['sdf'] - ov['sdf']).max()}"
    assert gv.tobytes() == ov.tobytes()
    assert st.voxels_updated == oc["total_updated"]
    assert st.blocks_visited == oc["total_touched"]
    assert st.frames_integrated == oc["frames_done"] and st.frames_skipped == oc["frames_skipped"]


@pytest.mark.parametrize("batch", [1, 3, 8])
def test_small_frames_color(built, batch):
    p = tsdf.default_params(width=160, height=120, max_blocks=1 << 15, hash_slots=1 << 17, batch_frames=batch)
    D, C, P, K = synth.make_frames(7, seed=3, width=160, height=120, loop_frames=200, noise_mm=1.5, drop=0.02)
    assert_identical(*run_both(p, D, C, P, K))


def test_single_frame_api_equals_batch(built):
    p = tsdf.default_params(width=160, height=120, max_blocks=1 << 15, hash_slots=1 << 17, batch_frames=4)
    D, C, P, K = synth.make_frames(5, seed=4, width=160, height=120, loop_frames=100)
    assert_identical(*run_both(p, D, None, P, K, batch_api=False))


def test_invalid_pose_and_ragged(built):
    """-inf poses are skipped (sensorData.h:382); width not a multiple of the 8x4 warp tile."""
    p = tsdf.default_params(width=150, height=97, max_blocks=1 << 17, hash_slots=1 << 19, batch_frames=4)
    D, C, P, K = synth.make_frames(9, seed=6, width=150, height=97, loop_frames=90, invalid_pose_every=3, drop=0.1)
    g, o = run_both(p, D, C, P, K)
    assert_identical(g, o)
    assert g[2].frames_skipped == 3


def test_empty_and_all_invalid_depth(built):
    p = tsdf.default_params(width=64, height=48, max_blocks=1 << 10, hash_slots=1 << 12, batch_frames=2)
    D = np.zeros((2, 48, 64), np.uint16); D[1] = 65535                       # 0 = invalid, 65.5 m > depth_max
    P = np.tile(np.eye(4, dtype=np.float32), (2, 1, 1)); K = synth.BoxRoomScene(width=64, height=48).intrinsics()
    g, o = run_both(p, D, None, P, K)
    assert_identical(g, o)
    assert len(g[0]) == 0


def test_weight_sample_and_saturation(built):
    """non-constant per-sample weight (weight_sample=10) and the u8 weight cap."""
    p = tsdf.default_params(width=96, height=72, max_blocks=1 << 14, hash_slots=1 << 16, batch_frames=8,
                            weight_sample=10, weight_max=40)
    D, C, P, K = synth.make_frames(12, seed=8, width=96, height=72, loop_frames=4000)
    g, o = run_both(p, D, C, P, K)
    assert_identical(g, o)
    assert g[1]["w"].max() == 40


def test_full_resolution_frames(built):
    """BASELINE.json sizes: 640x480, 4 mm voxels."""
    p = tsdf.default_params(max_blocks=1 << 16, hash_slots=1 << 18, batch_frames=4)
    D, C, P, K = synth.make_frames(4, seed=1, loop_frames=1000)
    g, o = run_both(p, D, C, P, K)
    assert_identical(g, o)
    # geometric sanity (size-independent property): |sdf| <= trunc at the observed depth range
    assert np.abs(g[1]["sdf"]).max() <= 0.02 + 0.01 * 6.0 + 1e-6


@pytest.mark.timeout(180)
@pytest.mark.parametrize("batch,flag", [(1, 0), (2, 0), (5, tsdf.KERNEL_TMA), (3, tsdf.KERNEL_COLUMN)])
def test_tma_staged_kernel_variant(built, batch, flag):
    """cp.async.bulk staged kernel (default for batches of <= 2 frames, forced with SCN_TSDF_KERNEL_TMA) — same bits"""
    p = tsdf.default_params(width=160, height=120, max_blocks=1 << 15, hash_slots=1 << 17, batch_frames=batch, flags=flag)
    D, C, P, K = synth.make_frames(7, seed=12, width=160, height=120, loop_frames=150, noise_mm=1.0, drop=0.03)
    assert_identical(*run_both(p, D, C, P, K))
    assert_identical(*run_both(p, D, None, P, K, batch_api=False))


def test_capacity_error_is_reported(built):
    p = tsdf.default_params(width=160, height=120, max_blocks=64, hash_slots=256, batch_frames=1)
    D, C, P, K = synth.make_frames(1, seed=3, width=160, height=120)
    vol = tsdf.TsdfVolume(p, device=0)
    vol.integrate_batch(D, None, P, K)
    from scannet_b200 import ScnError
    with pytest.raises(ScnError):
        vol.sync()
    vol.close()


def test_gpu_matches_committed_spec_hashes(built):
    """same seeded cases as tests/golden/tsdf_spec_golden.json, hashed from the CUDA path's own output"""
    import hashlib
    import importlib.util
    import json
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("mk", os.path.join(root, "scripts", "make_tsdf_golden.py"))
    mk = importlib.util.module_from_spec(spec); spec.loader.exec_module(mk)
    with open(os.path.join(root, "tests", "golden", "tsdf_spec_golden.json")) as fh:
        gold = json.load(fh)
    for c in gold["cases"]:
        w, h = c["wh"]
        D, C, P, K = synth.make_frames(c["frames"], seed=c["seed"], width=w, height=h, loop_frames=c["loop"], noise_mm=c["noise"],
                                       drop=c["drop"], invalid_pose_every=c["inv"])
        p = tsdf.default_params(width=w, height=h, max_blocks=1 << 15, hash_slots=1 << 17, batch_frames=4, **c["over"])
        vol = tsdf.TsdfVolume(p, device=0)
        vol.integrate_batch(D, C if c["color"] else None, P, K); vol.sync()
        xyz, vox = vol.download_blocks()
        mx, mc_, mt = vol.extract_mesh()
        assert hashlib.sha256(xyz.tobytes() + vox.tobytes()).hexdigest() == c["volume_sha256"], c["name"]
        assert hashlib.sha256(mx.tobytes() + mc_.tobytes() + mt.tobytes()).hexdigest() == c["mesh_sha256"], c["name"]
        vol.close()
