#!/bin/bash
# round 2, call Y: the shipped code of the round: suite, smoke, full bench line, reference arm
TAG=${1:-r02y}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > $OUT/gpu_$TAG.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -8 > $OUT/pytest_gpu_$TAG.log; tail -3 $OUT/pytest_gpu_$TAG.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke_$TAG.log 2>&1; tail -4 $OUT/smoke_$TAG.log
( time timeout 1500 python bench.py --steps 20 --warmup 5 > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err ) 2>&1 | tail -3; tail -c 300 $OUT/bench_$TAG.json; tail -3 $OUT/bench_$TAG.err
( time timeout 900 python bench.py --impl reference --steps 20 --warmup 5 > $OUT/bench_ref_$TAG.json 2> $OUT/bench_ref_$TAG.err ) 2>&1 | tail -3
