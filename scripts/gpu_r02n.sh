#!/bin/bash
# round 2, call N: k_alloc diet (fused key/mask map with a one-load fast path, branch-free DDA step, filtered-input template): parity + A/B
TAG=${1:-r02n}
OUT=gpurun_out
mkdir -p $OUT
timeout 1200 python -m pytest tests/test_tsdf_gpu.py tests/test_tsdf_bench_config_gpu.py tests/test_pipeline_gpu.py -q 2>&1 | tail -5
for v in prealloc cur prealloc cur; do
  lib=""; [ $v != cur ] && lib=$PWD/build/ab/$v/libscannet_b200.so
  echo "== $v"; SCN_B200_LIB=$lib timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu --no-seg 2>/dev/null | python -c "import sys,json; j=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=j['roofline']; print(j['value'], j['e2e']['value'], j['ms_per_step'], 'alloc_ms', r['alloc_kernel_ms_total'], 'int_ms', r['integrate_kernel_ms_total'])"
done 2>&1 | tee $OUT/ab_alloc_$TAG.txt
