#!/bin/bash
# round 2, call F: decoder rework (inflate window selection, JPEG cooperative bit reader), sort hybrid, tests + probe + full bench
TAG=${1:-r02f}
OUT=gpurun_out
mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -60 > $OUT/pytest_gpu_$TAG.log
tail -12 $OUT/pytest_gpu_$TAG.log
timeout 900 python scripts/probes/decode_probe.py > $OUT/decode_probe_$TAG.json 2> $OUT/decode_probe_$TAG.err; cat $OUT/decode_probe_$TAG.json | tr -d '\n ' | cut -c1-3000; echo; tail -3 $OUT/decode_probe_$TAG.err
( time timeout 1500 python bench.py --steps 20 --warmup 5 > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err ) 2>&1 | tail -3; tail -c 300 $OUT/bench_$TAG.json; tail -5 $OUT/bench_$TAG.err
