"""Dev probe (GPU): writes a synthetic .sens (zlib depth + JPEG colour) and runs bin/fuse on it with SCN_TIMING=1, both decode modes."""
import os, subprocess, sys, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np, cv2
from scannet_b200 import synth
n = int(sys.argv[1]) if len(sys.argv) > 1 else 600
D, C, P, K = synth.make_frames(n, seed=4, loop_frames=1500, noise_mm=1.0, drop=0.01)
rng = np.random.default_rng(0)
C = np.clip(C.astype(np.int16) + rng.integers(-12, 12, C.shape, dtype=np.int16), 0, 255).astype(np.uint8)
d = tempfile.mkdtemp(); p = os.path.join(d, "scene.sens")
synth.write_sens(p, D, C, P, K, depth_comp=1, color_comp=2, jpeg_encoder=lambda x: cv2.imencode(".jpg", x[:, :, ::-1], [int(cv2.IMWRITE_JPEG_QUALITY), 85])[1].tobytes())
prm = os.path.join(d, "p.txt"); open(prm, "w").write("s_SDFVoxelSize = 0.004f;\ns_SDFTruncation = 0.02f;\ns_SDFTruncationScale = 0.01f;\ns_hashNumSDFBlocks = 3000000;\n")
for mode in ("gpu", "host", "gpu", "host"):
    r = subprocess.run([os.path.join(ROOT, "scannet_b200", "bin", "fuse"), prm, p, os.path.join(d, "o.ply")], capture_output=True, text=True, env=dict(os.environ, SCN_FUSE_DECODE=mode, SCN_TIMING="1"))
    print(mode, [l for l in r.stdout.splitlines() if l.startswith("integrated")][0][:90]); print(r.stderr.strip()[-300:])
