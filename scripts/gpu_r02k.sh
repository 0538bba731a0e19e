#!/bin/bash
# round 2, call K: inflate with the 8-bit distance table (tests, probe), ncu source counters of the JPEG entropy kernel
TAG=${1:-r02k}
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m pytest tests/test_inflate_gpu.py -q 2>&1 | tail -4
timeout 600 python scripts/probes/decode_probe.py inflate > $OUT/decode_probe_$TAG.json 2> $OUT/decode_probe_$TAG.err; python - <<PY
import json
j=json.load(open('$OUT/decode_probe_$TAG.json'))
for k,v in j['inflate'].items(): print(k,v)
PY
timeout 900 ncu --set full --import-source on --clock-control none -k regex:k_jpeg_entropy -c 1 -o $OUT/ncu_jpeg_$TAG -f python scripts/probes/jpeg_one.py > $OUT/ncu_jpeg_$TAG.log 2>&1; tail -3 $OUT/ncu_jpeg_$TAG.log
