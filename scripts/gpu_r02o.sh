#!/bin/bash
# round 2, call O: k_alloc follow-ups — sentinel hoisted, 6 CTAs/SM, group sizes; parity first
TAG=${1:-r02o}
OUT=gpurun_out
mkdir -p $OUT
timeout 1200 python -m pytest tests/test_tsdf_gpu.py tests/test_tsdf_bench_config_gpu.py -q 2>&1 | tail -3
run() { SCN_B200_LIB=$1 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu --no-seg 2>/dev/null | python -c "import sys,json; j=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=j['roofline']; print(j['value'], j['e2e']['value'], j['ms_per_step'], 'alloc_ms', r['alloc_kernel_ms_total'], 'int_ms', r['integrate_kernel_ms_total'])"; }
{
echo "== cur"; run ""
echo "== alloc6"; run $PWD/build/ab/alloc6/libscannet_b200.so
echo "== cur group8"; SCN_TSDF_ALLOC_GROUP=8 run ""
echo "== cur group2"; SCN_TSDF_ALLOC_GROUP=2 run ""
echo "== alloc6 group8"; SCN_TSDF_ALLOC_GROUP=8 run $PWD/build/ab/alloc6/libscannet_b200.so
echo "== cur reserve2"; SCN_TSDF_RESERVE=2 run ""
echo "== cur"; run ""
} 2>&1 | tee $OUT/ab_alloc_$TAG.txt
