"""CPU: mutation fuzzing of the host-side parsers under AddressSanitizer + UBSan (tests/fuzz/host_fuzz.cpp, built from the
product sources): .sens container + inflate + baseline / progressive JPEG + PNG, PLY / OBJ, segs.json.  Thousands of corrupted
inputs must all come back as status codes — no crash, no sanitizer report, no hang."""
import json
import os
import shutil
import subprocess

import numpy as np
import pytest

from scannet_b200 import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CXX = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
FLAGS = ["-O1", "-g", "-std=c++17", "-fsanitize=address,undefined", "-fno-sanitize=shift", "-fno-sanitize-recover=undefined", "-fno-omit-frame-pointer"]
# -fno-sanitize=shift: the IDCT shifts negative ints left exactly as stb_image does (arithmetic shift on every supported compiler)


@pytest.fixture(scope="module")
def fuzzer(tmp_path_factory):
    d = tmp_path_factory.mktemp("fuzz")
    exe = str(d / "host_fuzz")
    srcs = [os.path.join(ROOT, "tests", "fuzz", "host_fuzz.cpp")] + [os.path.join(ROOT, "scannet_b200", "csrc", f) for f in ("jpeg.cpp", "sens.cpp", "mesh_io.cpp")]
    r = subprocess.run([CXX] + FLAGS + ["-o", exe] + srcs + ["-lpthread"], capture_output=True, text=True)
    if r.returncode != 0 and "asan" in (r.stderr or "").lower():
        pytest.skip("AddressSanitizer runtime not available")
    assert r.returncode == 0, r.stderr[-2000:]
    return exe, d


def run(exe, kind, seed, work, iters):
    r = subprocess.run([exe, kind, str(seed), str(work), str(iters)], capture_output=True, text=True, timeout=600,
                       env=dict(os.environ, ASAN_OPTIONS="detect_leaks=0:allocator_may_return_null=0"))
    assert r.returncode == 0, (r.stdout[-500:], r.stderr[-3000:])
    assert "runtime error" not in r.stderr and "AddressSanitizer" not in r.stderr, r.stderr[-3000:]
    return r.stdout


@pytest.mark.parametrize("colour", ["jpeg", "progressive", "png", "raw"])
def test_sens_container_and_decoders(fuzzer, colour):
    import cv2
    exe, d = fuzzer
    D, C, P, K = synth.make_frames(2, seed=3, width=64, height=48, loop_frames=20, noise_mm=1.0)
    jpg = lambda prog: (lambda x: cv2.imencode(".jpg", x[:, :, ::-1], [int(cv2.IMWRITE_JPEG_QUALITY), 80, int(cv2.IMWRITE_JPEG_PROGRESSIVE), prog])[1].tobytes())
    seed = str(d / f"seed_{colour}.sens")
    if colour == "raw": synth.write_sens(seed, D, C, P, K, depth_comp=0, color_comp=0)
    elif colour == "png": synth.write_sens(seed, D, C, P, K, depth_comp=1, color_comp=1, jpeg_encoder=lambda x: cv2.imencode(".png", x[:, :, ::-1])[1].tobytes())
    else: synth.write_sens(seed, D, C, P, K, depth_comp=1, color_comp=2, jpeg_encoder=jpg(1 if colour == "progressive" else 0))
    out = run(exe, "sens", seed, d, 4000)
    assert "4000 inputs" in out


@pytest.mark.parametrize("fmt", ["ply_binary", "ply_ascii", "obj"])
def test_mesh_loaders(fuzzer, fmt):
    exe, d = fuzzer
    xyz, tri = synth.make_grid_mesh(20, 15, 1)
    if fmt == "ply_binary":
        seed = os.path.join(ROOT, "tests", "golden", "gates381.ply"); iters = 600
    elif fmt == "ply_ascii":
        seed = str(d / "a.ply"); iters = 2000
        with open(seed, "w") as f:
            f.write(f"ply\nformat ascii 1.0\nelement vertex {len(xyz)}\nproperty float x\nproperty float y\nproperty float z\nelement face {len(tri)}\nproperty list uchar int vertex_indices\nend_header\n")
            for v in xyz: f.write(f"{v[0]:.6f} {v[1]:.6f} {v[2]:.6f}\n")
            for t in tri: f.write(f"3 {t[0]} {t[1]} {t[2]}\n")
    else:
        seed = str(d / "m.obj"); iters = 2000
        with open(seed, "w") as f:
            for v in xyz: f.write(f"v {v[0]:.6f} {v[1]:.6f} {v[2]:.6f}\n")
            for t in tri: f.write(f"f {t[0] + 1} {t[1] + 1} {t[2] + 1}\n")
    assert f"{iters} inputs" in run(exe, "mesh", seed, d, iters)


def test_segs_json_reader(fuzzer):
    exe, d = fuzzer
    seed = str(d / "s.segs.json")
    json.dump({"params": {"kThresh": 0.01, "segMinVerts": 20}, "sceneId": "/x", "segIndices": list(range(0, 3000, 7)) * 3}, open(seed, "w"), separators=(",", ":"))
    assert "4000 inputs" in run(exe, "segs", seed, d, 4000)


def test_read_ahead_cache_under_thread_sanitizer(tmp_path):
    """the RGBDFrameCacheRead counterpart (scn_sens_cache_*) built with -fsanitize=thread: producers, consumer and early destruction"""
    import cv2
    exe = str(tmp_path / "host_tsan")
    srcs = [os.path.join(ROOT, "tests", "fuzz", "host_fuzz.cpp")] + [os.path.join(ROOT, "scannet_b200", "csrc", f) for f in ("jpeg.cpp", "sens.cpp", "mesh_io.cpp")]
    r = subprocess.run([CXX, "-O1", "-g", "-std=c++17", "-fsanitize=thread", "-DSCN_NO_TARGET_CLONES", "-o", exe] + srcs + ["-lpthread"], capture_output=True, text=True)
    if r.returncode != 0 and "tsan" in (r.stderr or "").lower():
        pytest.skip("ThreadSanitizer runtime not available")
    assert r.returncode == 0, r.stderr[-2000:]
    D, C, P, K = synth.make_frames(12, seed=3, width=64, height=48, loop_frames=40, noise_mm=1.0)
    seed = str(tmp_path / "c.sens")
    synth.write_sens(seed, D, C, P, K, depth_comp=1, color_comp=2, jpeg_encoder=lambda x: cv2.imencode(".jpg", x[:, :, ::-1])[1].tobytes())
    r = subprocess.run([exe, "cache", seed], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "ThreadSanitizer" not in r.stderr, (r.stdout[-300:], r.stderr[-3000:])
    assert "frames delivered in order" in r.stdout
