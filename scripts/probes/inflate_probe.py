"""Dev probe (GPU): inflate N synthetic 640x480 depth frames on the device; prints frames/s.  Used under ncu for k_inflate."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np, torch
from scannet_b200 import sens, synth
from scannet_b200.sens import SensFile
import tempfile
N = int(sys.argv[1]) if len(sys.argv) > 1 else 64
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
sc = synth.BoxRoomScene(seed=3)
P = np.stack([sc.camera_pose(i, 1000) for i in range(N)])
D = np.stack([sc.render(P[i], noise_mm=1.0, frame_seed=i)[0] for i in range(N)])
d = tempfile.mkdtemp(); p = os.path.join(d, "s.sens")
synth.write_sens(p, D, None, P, sc.intrinsics(), depth_comp=1, color_comp=0)
f = SensFile(p)
out = torch.zeros((N, 480, 640), dtype=torch.int16, device="cuda")
f.decode_depth_device(0, min(N, 8), out.data_ptr())
torch.cuda.synchronize()
for _ in range(reps):
    t0 = time.perf_counter(); f.decode_depth_device(0, N, out.data_ptr()); dt = time.perf_counter() - t0
    print(N, "frames", round(dt * 1e3, 1), "ms", round(N / dt), "fps")
assert (out.cpu().numpy().view(np.uint16) == D).all()
