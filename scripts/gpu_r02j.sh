#!/bin/bash
# round 2, call J: MCU-batched JPEG IDCT + pooled staging + lean inflate: full GPU suite, decode probe, full bench
TAG=${1:-r02j}
OUT=gpurun_out
mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -40 > $OUT/pytest_gpu_$TAG.log; tail -6 $OUT/pytest_gpu_$TAG.log
timeout 900 python scripts/probes/decode_probe.py > $OUT/decode_probe_$TAG.json 2> $OUT/decode_probe_$TAG.err; python - <<PY
import json
j=json.load(open('$OUT/decode_probe_$TAG.json'))
for k,v in j['jpeg'].items(): print(k,v)
PY
tail -3 $OUT/decode_probe_$TAG.err
( time timeout 1500 python bench.py --steps 20 --warmup 5 > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err ) 2>&1 | tail -3; tail -c 300 $OUT/bench_$TAG.json; tail -5 $OUT/bench_$TAG.err
