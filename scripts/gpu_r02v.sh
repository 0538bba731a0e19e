#!/bin/bash
# round 2, call V: tier-2 sort in shared memory (one warp per segment): parity + timings at C1 / C5
TAG=${1:-r02v}
OUT=gpurun_out
mkdir -p $OUT
timeout 1200 python -m pytest tests/test_segmentator_gpu.py tests/test_segs_gpu.py -q 2>&1 | tail -3
python scripts/probes/sort_phase_probe.py
SCN_SEG_SORT_LAUNCHES=1 python scripts/probes/sort_phase_probe.py
timeout 900 python - <<'PY'
import sys, json
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import bench
out = bench.seg_bench(True)
for k in ("c1_50k", "c5_2m"):
    v = out[k]; print(k, v["stages_ms"], v["sort_kernel_launches"], "ref_s", round(v["cpu_reference_s"], 4), "identical", v["bit_identical_to_reference"], v["bit_identical_to_cpu"])
json.dump(out, open("gpurun_out/seg_bench_r02v.json", "w"))
PY
