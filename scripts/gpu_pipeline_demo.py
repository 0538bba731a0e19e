#!/usr/bin/env python
"""End-to-end demonstration at BASELINE.json configs[2] scale (stand-in for scene0000_00, which is licence-gated):
synthetic .sens (640x480 zlib depth + JPEG colour) -> bin/fuse (TSDF + marching cubes) -> bin/segmentator.
Prints one JSON object.  Usage: python scripts/gpu_pipeline_demo.py [n_frames]"""
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
from scannet_b200 import synth  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 600
    import cv2
    out = {"frames": n}
    with tempfile.TemporaryDirectory() as d:
        t0 = time.perf_counter()
        D, C, P, K = synth.make_frames(n, seed=4, loop_frames=1500, noise_mm=1.0, drop=0.01, invalid_pose_every=97)
        rng = np.random.default_rng(0)
        C = np.clip(C.astype(np.int16) + rng.integers(-12, 12, C.shape, dtype=np.int16), 0, 255).astype(np.uint8)
        p = os.path.join(d, "scene.sens")
        synth.write_sens(p, D, C, P, K, depth_comp=1, color_comp=2,
                         jpeg_encoder=lambda x: cv2.imencode(".jpg", x[:, :, ::-1], [int(cv2.IMWRITE_JPEG_QUALITY), 85])[1].tobytes())
        out["generate_s"] = time.perf_counter() - t0
        out["sens_mb"] = os.path.getsize(p) / 1e6
        prm = os.path.join(d, "zParameters.txt")
        with open(prm, "w") as fh:
            fh.write("s_SDFVoxelSize = 0.004f;\ns_SDFTruncation = 0.02f;\ns_SDFTruncationScale = 0.01f;\ns_hashNumSDFBlocks = 3000000;\n")
        # depth inflated on the GPU from the compressed payloads (whole scan resident in HBM), colour on the host pool
        t0 = time.perf_counter()
        rg = subprocess.run([os.path.join(ROOT, "scannet_b200", "bin", "fuse"), prm, p, os.path.join(d, "gpu_decode.ply")], capture_output=True, text=True,
                            env=dict(os.environ, SCN_FUSE_DECODE="gpu"))
        out["fuse_gpu_decode_wall_s"] = time.perf_counter() - t0
        out["fuse_gpu_decode_stdout"] = rg.stdout.strip().splitlines()[-4:]
        t0 = time.perf_counter()
        r = subprocess.run([os.path.join(ROOT, "scannet_b200", "bin", "fuse"), prm, p], capture_output=True, text=True)
        out["fuse_wall_s"] = time.perf_counter() - t0
        out["fuse_rc"] = r.returncode
        out["fuse_stdout"] = r.stdout.strip().splitlines()[-4:]
        if r.returncode:
            out["fuse_stderr"] = r.stderr[-400:]
        try:
            out["gpu_decode_ply_identical"] = open(os.path.join(d, "gpu_decode.ply"), "rb").read() == open(os.path.join(d, "scene_vh.ply"), "rb").read()
        except OSError:
            out["gpu_decode_ply_identical"] = None
        ply = os.path.join(d, "scene_vh.ply")
        if os.path.exists(ply):
            out["ply_mb"] = os.path.getsize(ply) / 1e6
            t0 = time.perf_counter()
            r2 = subprocess.run([os.path.join(ROOT, "scannet_b200", "bin", "segmentator"), ply], capture_output=True, text=True)
            out["segmentator_wall_s"] = time.perf_counter() - t0
            out["segmentator_stdout"] = r2.stdout.strip().splitlines()[-2:]
            ref = os.path.join(ROOT, "oracle", "_ref", "segmentator_ref_O2")
            if os.path.exists(ref) and os.environ.get("WITH_REF", "1") == "1":
                seg_ours = open(os.path.join(d, "scene_vh.0.010000.segs.json"), "rb").read()
                t0 = time.perf_counter()
                subprocess.run([ref, ply], capture_output=True, text=True)
                out["reference_segmentator_wall_s"] = time.perf_counter() - t0
                out["segs_json_identical_to_reference"] = seg_ours == open(os.path.join(d, "scene_vh.0.010000.segs.json"), "rb").read()
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
