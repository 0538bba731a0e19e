"""Where the time of one file -> TSDF scene goes (scn_fuse_scene reports) against the decode chunk size.
Usage: python scripts/probes/fuse_probe.py [frames] > gpurun_out/fuse_probe.json"""
import json
import os
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import bench  # noqa: E402
from scannet_b200 import fuse as sfuse  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
dev = torch.device("cuda:0")
out = {}
with tempfile.TemporaryDirectory() as d:
    p = os.path.join(d, "scene.sens")
    bench.make_sens_file(p, n, 100, dev, with_color=False)
    for chunk in ("512", "1024", "2048", "256"):
        os.environ["SCN_FUSE_CHUNK"] = chunk
        runs = []
        for rep_i in range(3):
            t0 = time.perf_counter()
            rep = sfuse.fuse_scene(p, None, decode_mode="gpu", device=0)
            rep["wall_s"] = time.perf_counter() - t0
            runs.append({k: (round(v, 4) if isinstance(v, float) else v) for k, v in rep.items()})
        out["chunk_" + chunk] = runs
print(json.dumps(out, indent=1))
