#!/bin/bash
# round 2, call L: where the rest of fuse_s goes at 5,578 frames (buffer allocation / teardown timers)
TAG=${1:-r02l}
OUT=gpurun_out
mkdir -p $OUT
timeout 1500 python scripts/probes/fuse_probe.py 5578 1 quick > $OUT/fuse_probe_c3full_$TAG.json 2> $OUT/fuse_probe_c3full_$TAG.err; tail -2 $OUT/fuse_probe_c3full_$TAG.err
python - <<PY
import json
j=json.load(open('$OUT/fuse_probe_c3full_$TAG.json'))
for k,v in j.items():
    for r in v: print(k, r)
PY
