"""One launch of each GPU inflate variant over 592 depth streams (for ncu)."""
import os, sys, zlib
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch
from scannet_b200 import sens, synth
sc = synth.BoxRoomScene(size=(6.0, 5.0, 3.0), seed=3, width=640, height=480)
Z = [zlib.compress(sc.render(sc.camera_pose(i, 1000), noise_mm=1.0, frame_seed=i)[0].tobytes(), 6) for i in range(16)]
n = 592
streams = (Z * (n // len(Z) + 1))[:n]
dout = torch.empty((n, 480, 640), dtype=torch.int16, device="cuda")
for window in ("ring", "hbm"):
    os.environ["SCN_INFLATE_WINDOW"] = window
    sens.inflate_batch_device(streams, 640 * 480 * 2, dout.data_ptr())
torch.cuda.synchronize()
