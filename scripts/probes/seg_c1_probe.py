"""Dev probe (GPU): per-stage Segmentator timings over repeated calls.  args: [seed] [torch]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
seed = int(sys.argv[1]) if len(sys.argv) > 1 else 0
if len(sys.argv) > 2:
    import torch
    torch.zeros(1, device="cuda"); print("torch threads", torch.get_num_threads())
from scannet_b200 import segmentator, synth
for n in ((250, 200), (1600, 1250)):
    xyz, tri = synth.make_feature_mesh(n[0], n[1], seed)
    for i in range(4):
        t0 = time.perf_counter(); segmentator.segment_mesh(xyz, tri); dt = time.perf_counter() - t0
        ms, _ = segmentator.last_timings()
        print(n, "seed", seed, i, round(dt * 1e3, 2), [round(x, 2) for x in ms])
