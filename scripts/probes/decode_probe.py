"""Kernel-only decode rates (CUDA events inside the library) of the GPU inflate (both window placements) and the GPU JPEG decoder
against the number of streams in one launch.  Usage: python scripts/probes/decode_probe.py > gpurun_out/decode_probe.json"""
import json
import os
import sys
import zlib

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from scannet_b200 import sens, synth  # noqa: E402

out = {"inflate": {}, "jpeg": {}}
sc = synth.BoxRoomScene(size=(6.0, 5.0, 3.0), seed=3, width=640, height=480)
D = [sc.render(sc.camera_pose(i, 1000), noise_mm=1.0, frame_seed=i)[0] for i in range(48)]
Z = [zlib.compress(d.tobytes(), 6) for d in D]
for window in ("ring", "hbm"):
    os.environ["SCN_INFLATE_WINDOW"] = window
    for n in (96, 240, 480, 740, 960, 1480, 1920, 3840):
        streams = (Z * (n // len(Z) + 1))[:n]
        dout = torch.empty((n, 480, 640), dtype=torch.int16, device="cuda")
        sens.inflate_batch_device(streams[:48], 640 * 480 * 2, dout.data_ptr())
        sens.inflate_batch_device(streams, 640 * 480 * 2, dout.data_ptr())
        pk, kms, ring, _ = sens.inflate_last_timings()
        ok = bool((dout[n - 1].cpu().numpy().view(np.uint16) == D[(n - 1) % 48]).all())
        out["inflate"][f"{window}_{n}"] = {"kernel_ms": round(kms, 3), "kernel_fps": round(n / (kms * 1e-3)), "pack_ms": round(pk * 1e3, 2), "ok": ok}
        del dout
os.environ.pop("SCN_INFLATE_WINDOW")
try:
    if len(sys.argv) > 1 and sys.argv[1] == "inflate":
        raise RuntimeError("inflate only")
    import cv2
    for (w, h) in ((640, 480), (1296, 968)):
        yy, xx = np.mgrid[0:h, 0:w]
        J = []
        for i in range(24):
            im = np.stack([(xx * 255 // w + 3 * i) % 256, (yy * 255 // h + 5 * i) % 256, ((xx + yy) // 3 + 7 * i) % 256], -1).astype(np.uint8)
            im = cv2.GaussianBlur(im, (0, 0), 1.5)
            J.append(cv2.imencode(".jpg", im, [int(cv2.IMWRITE_JPEG_QUALITY), 85])[1].tobytes())
        for n in (96, 480, 960, 1920, 2880):
            jp = (J * (n // 24 + 1))[:n]
            dout = torch.empty((n, h, w, 3), dtype=torch.uint8, device="cuda")
            sens.jpeg_decode_batch_device(jp[:24], w, h, dout.data_ptr())
            k = sens.jpeg_decode_batch_device(jp, w, h, dout.data_ptr())
            hs, ems, cms = sens.jpeg_last_timings()
            out["jpeg"][f"{w}x{h}_{n}"] = {"entropy_idct_ms": round(ems, 3), "colour_ms": round(cms, 3), "kernels_fps": round(n / ((ems + cms) * 1e-3)), "host_ms": round(hs * 1e3, 2), "on_device": k}
            del dout
except Exception as e:
    out["jpeg"]["error"] = repr(e)
print(json.dumps(out, indent=1))
