#!/bin/bash
# One GPU session: parity tests, smoke, bench (both arms), ncu launch list + one full capture of the top kernel.
# Usage (from the repo root on the GPU box): bash scripts/gpu_check.sh [tag]
TAG=${1:-r01}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > $OUT/gpu_$TAG.txt 2>&1
nproc >> $OUT/gpu_$TAG.txt; lscpu | grep "Model name" >> $OUT/gpu_$TAG.txt
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > $OUT/pytest_gpu_$TAG.log
cat $OUT/pytest_gpu_$TAG.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke_$TAG.log 2>&1; tail -5 $OUT/smoke_$TAG.log
timeout 600 python bench.py --impl reference > $OUT/bench_ref_$TAG.json 2> $OUT/bench_ref_$TAG.err; tail -c 600 $OUT/bench_ref_$TAG.json
timeout 900 python bench.py --seg-c5 > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err; cat $OUT/bench_$TAG.json; tail -5 $OUT/bench_$TAG.err
for B in 1 16; do timeout 600 python bench.py --no-cpu --no-seg --batch $B > $OUT/bench_batch${B}_$TAG.json 2>/dev/null; python -c "
import json,sys; d=json.load(open('$OUT/bench_batch${B}_$TAG.json')); print('batch $B', round(d['value']), round(d['e2e']['value']), d['roofline']['avg_launch_ms'], d['roofline']['alloc_kernel_ms_total'], d['roofline']['integrate_kernel_ms_total'])"; done
if [ "${NCU:-1}" = "1" ]; then
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_ -c 400 --csv --log-file $OUT/launches_$TAG.csv \
      python bench.py --steps 2 --warmup 1 --frames-per-step 48 --no-cpu --no-seg > $OUT/ncu_launch_$TAG.log 2>&1
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_integrate -s 4 -c 2 -f -o $OUT/prof_integrate_$TAG \
      python bench.py --steps 2 --warmup 1 --frames-per-step 48 --no-cpu --no-seg > $OUT/ncu_full_$TAG.log 2>&1
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_alloc -s 4 -c 1 -f -o $OUT/prof_alloc_$TAG \
      python bench.py --steps 2 --warmup 1 --frames-per-step 48 --no-cpu --no-seg > $OUT/ncu_full_alloc_$TAG.log 2>&1
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_inflate -s 1 -c 1 -f -o $OUT/prof_inflate_$TAG \
      python scripts/probes/inflate_probe.py 64 1 > $OUT/ncu_full_inflate_$TAG.log 2>&1
  tail -3 $OUT/ncu_full_$TAG.log
fi
