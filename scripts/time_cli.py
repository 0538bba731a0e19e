#!/usr/bin/env python
"""Times bin/segmentator (with its SCN_TIMING stage report) against the reference CLI on the C5 mesh."""
import os, subprocess, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from scannet_b200 import synth
x, t = synth.make_feature_mesh(1600, 1250, 5); synth.write_ply("/tmp/c5.ply", x, t)
for i in range(2):
    t0 = time.perf_counter()
    r = subprocess.run([os.path.join(ROOT, "scannet_b200/bin/segmentator"), "/tmp/c5.ply"], capture_output=True, text=True, env=dict(os.environ, SCN_TIMING="1"))
    print("ours wall %.3f s" % (time.perf_counter() - t0), r.stderr.strip())
ref = os.path.join(ROOT, "oracle/_ref/segmentator_ref_O2")
if os.path.exists(ref):
    t0 = time.perf_counter(); subprocess.run([ref, "/tmp/c5.ply"], capture_output=True); print("reference wall %.3f s" % (time.perf_counter() - t0))
