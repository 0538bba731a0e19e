"""Does a 1000-frame scene integrate as fast from a cold start (fresh volume, idle GPU) as in the steady state of bench.py?
Times scn_tsdf_integrate_device + sync (host clock) on the decoded depth of one synthetic .sens file.
Usage: python scripts/probes/integrate_idle_probe.py > gpurun_out/integrate_idle.json"""
import json
import os
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import bench  # noqa: E402
from scannet_b200 import sens, tsdf  # noqa: E402

n = 1000
dev = torch.device("cuda:0")
out = {}
with tempfile.TemporaryDirectory() as d:
    p = os.path.join(d, "scene.sens")
    bench.make_sens_file(p, n, 100, dev, with_color=False)
    sf = sens.SensFile(p)
    dd = torch.empty((n, 480, 640), dtype=torch.int16, device=dev)
    sf.decode_depth_device(0, n, dd.data_ptr())
    torch.cuda.synchronize()
    P = np.stack([sf.pose(i) for i in range(n)])
    K = sf.K_depth()

    def run(vol):
        t0 = time.perf_counter()
        vol.integrate_device(n, dd.data_ptr(), None, P, K)
        vol.sync()
        return round((time.perf_counter() - t0) * 1e3, 2)

    res = []
    for i in range(3):
        t0 = time.perf_counter(); v = tsdf.TsdfVolume(); tc = round((time.perf_counter() - t0) * 1e3, 2)
        a = run(v); v.reset(); b = run(v); v.reset(); c = run(v)
        time.sleep(0.3); v.reset(); e = run(v)
        v.profile(True); v.reset(); f = run(v); kt = v.kernel_times()
        t0 = time.perf_counter(); v.close(); td = round((time.perf_counter() - t0) * 1e3, 2)
        res.append({"create_ms": tc, "fresh_volume_ms": a, "after_reset_ms": b, "after_reset2_ms": c, "after_0.3s_idle_ms": e, "profiled_ms": f,
                    "alloc_kernel_ms": round(kt[0], 2), "integrate_kernel_ms": round(kt[1], 2), "destroy_ms": td})
    out["runs"] = res
os.write(bench._REAL_STDOUT, (json.dumps(out, indent=1) + "\n").encode())
