#!/bin/bash
# round 2, call U: pooled frame buffers in the scene driver: pipeline tests, configs[2] probe (4 runs), sort phase probe
TAG=${1:-r02u}
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m pytest tests/test_pipeline_gpu.py tests/test_jpeg_gpu.py tests/test_inflate_gpu.py -q 2>&1 | tail -3
timeout 1500 python scripts/probes/fuse_probe.py 5578 1 quick > $OUT/fuse_probe_c3full_$TAG.json 2> $OUT/fuse_probe_c3full_$TAG.err; tail -2 $OUT/fuse_probe_c3full_$TAG.err
python - <<PY
import json
j=json.load(open('$OUT/fuse_probe_c3full_$TAG.json'))
for k,v in j.items():
    for r in v: print(k, r)
PY
bash scripts/gpu_r02t.sh $TAG
