#!/bin/bash
# round 2, call Q: k_alloc depth prefetch + packed colour update: parity, A/B (depth-only and colour), colour CTAs 12 vs 14
TAG=${1:-r02q}
OUT=gpurun_out
mkdir -p $OUT
timeout 1200 python -m pytest tests/test_tsdf_gpu.py tests/test_tsdf_bench_config_gpu.py tests/test_pipeline_gpu.py -q 2>&1 | tail -4
run() { SCN_B200_LIB=$1 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu --no-seg $2 2>/dev/null | python -c "import sys,json; j=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=j['roofline']; print(j['value'], j['e2e']['value'], j['ms_per_step'], 'alloc_ms', r['alloc_kernel_ms_total'], 'int_ms', r['integrate_kernel_ms_total'])"; }
{
echo "== prealloc (r02j kernels)"; run $PWD/build/ab/prealloc/libscannet_b200.so
echo "== cur"; run ""
echo "== cur"; run ""
echo "== prealloc colour"; run $PWD/build/ab/prealloc/libscannet_b200.so --color
echo "== cur colour (12 CTAs)"; run "" --color
echo "== colour 14 CTAs"; run $PWD/build/ab/color14/libscannet_b200.so --color
} 2>&1 | tee $OUT/ab_$TAG.txt
