"""One launch of the GPU JPEG decoder over 296 frames of 1296x968 4:2:0 (for ncu)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import cv2
import torch
from scannet_b200 import sens
w, h = 1296, 968
yy, xx = np.mgrid[0:h, 0:w]
J = []
for i in range(8):
    im = np.stack([(xx * 255 // w + 3 * i) % 256, (yy * 255 // h + 5 * i) % 256, ((xx + yy) // 3 + 7 * i) % 256], -1).astype(np.uint8)
    im = cv2.GaussianBlur(im, (0, 0), 1.5)
    J.append(cv2.imencode(".jpg", im, [int(cv2.IMWRITE_JPEG_QUALITY), 85])[1].tobytes())
n = 296
jp = (J * (n // 8 + 1))[:n]
dout = torch.empty((n, h, w, 3), dtype=torch.uint8, device="cuda")
sens.jpeg_decode_batch_device(jp, w, h, dout.data_ptr())
torch.cuda.synchronize()
print(len(J[0]), sens.jpeg_last_timings())
