#!/bin/bash
# round 2, call X: volume pooled per device: pipeline tests + file->TSDF rates (single scene, 8 scenes) three times
TAG=${1:-r02x}
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m pytest tests/test_pipeline_gpu.py tests/test_tsdf_gpu.py -q 2>&1 | tail -3
for i in 1 2 3; do timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu --no-seg-c5 --c3-frames 0 2>/dev/null | python -c "
import sys,json
j=json.loads(sys.stdin.read().strip().splitlines()[-1]); f=j['file_to_tsdf']
print('value', round(j['value']), 'f2t', round(f['value']), round(f['rank0']['frames_per_s_incl_decode']), 'reused', f['rank0']['volume_reused'], 'many', round(f['many']['value']), f['many']['wall_s_max_over_ranks'], f['many']['volume_reused'])"; done
