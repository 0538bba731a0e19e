"""Segmentator at C1 (50 k vertices) twice: warm-up + measured; prints the library's stage timings.  Run under
`ncu --metrics gpu__time_duration.sum` with SCN_SEG_SORT_LAUNCHES=1 to get the per-phase kernel times of the sort."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from scannet_b200 import segmentator, synth
xyz, tri = synth.make_feature_mesh(250, 200, seed=5)
segmentator.segment_mesh(xyz, tri)
segmentator.segment_mesh(xyz, tri)
ms, n = segmentator.last_timings()
print("stages_ms", [round(float(x), 3) for x in ms], "sort launches", n)
