#!/bin/bash
# round 2, call A: parity tests (incl. the new bench-instantiation tests), new bench line, reference arm
TAG=${1:-r02a}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > $OUT/gpu_$TAG.txt 2>&1
nproc >> $OUT/gpu_$TAG.txt; lscpu | grep -E "Model name|Socket|NUMA|Thread" >> $OUT/gpu_$TAG.txt; free -g >> $OUT/gpu_$TAG.txt
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > $OUT/pytest_gpu_$TAG.log
cat $OUT/pytest_gpu_$TAG.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke_$TAG.log 2>&1; tail -5 $OUT/smoke_$TAG.log
( time timeout 900 python bench.py --steps 20 --warmup 5 > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err ) 2>&1 | tail -3; cat $OUT/bench_$TAG.json; tail -5 $OUT/bench_$TAG.err
( time timeout 900 python bench.py --impl reference --steps 20 --warmup 5 > $OUT/bench_ref_$TAG.json 2> $OUT/bench_ref_$TAG.err ) 2>&1 | tail -3; tail -c 1500 $OUT/bench_ref_$TAG.json
