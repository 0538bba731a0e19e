"""Dev probe (CPU only): times host union-find replay variants on the 2 M-vertex synthetic mesh.
Edges come from the test oracle; build + run:  python scripts/probes/kruskal_probe.py"""
import os, subprocess, sys, tempfile
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_bindings as ob
from scannet_b200 import synth

d = tempfile.mkdtemp()
xyz, tri = synth.make_feature_mesh(1600, 1250, 0)
seg, pre, srt, roots, nrm = ob.oracle_segment(xyz, tri, want_debug=True)
a = srt["a"].astype(np.int64); b = srt["b"].astype(np.int64)
key = np.maximum(a, b) * (1 << 32) + np.minimum(a, b)
_, first = np.unique(key, return_index=True)
keep = np.zeros(len(srt), bool); keep[first] = True; keep &= a != b
srt[keep].tofile(os.path.join(d, "edges.bin"))
subprocess.check_call(["/usr/bin/g++", "-O3", "-w", "-o", os.path.join(d, "h"), os.path.join(ROOT, "scripts/probes/kruskal_probe.cpp")])
print(subprocess.run([os.path.join(d, "h")], cwd=d, capture_output=True, text=True).stdout)
