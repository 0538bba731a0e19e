#!/bin/bash
# round 2, call I: pooled decode staging (fuse breakdown again), cold-vs-hot integrate probe, ncu source counters of the new inflate loop
TAG=${1:-r02i}
OUT=gpurun_out
mkdir -p $OUT
timeout 600 python scripts/probes/integrate_idle_probe.py > $OUT/integrate_idle_$TAG.json 2> $OUT/integrate_idle_$TAG.err; cat $OUT/integrate_idle_$TAG.json | tr -d '\n' | cut -c1-2500; echo; tail -2 $OUT/integrate_idle_$TAG.err
timeout 900 python scripts/probes/fuse_probe.py 1000 0 > $OUT/fuse_probe_depth_$TAG.json 2> $OUT/fuse_probe_depth_$TAG.err; tail -2 $OUT/fuse_probe_depth_$TAG.err
timeout 1500 python scripts/probes/fuse_probe.py 2048 1 > $OUT/fuse_probe_c3_$TAG.json 2> $OUT/fuse_probe_c3_$TAG.err; tail -2 $OUT/fuse_probe_c3_$TAG.err
python - <<PY
import json
for f in ('depth','c3'):
    j=json.load(open('$OUT/fuse_probe_%s_$TAG.json' % f))
    for k,v in j.items():
        for r in v: print(f, k, r)
PY
timeout 900 ncu --set full --import-source on --clock-control none -k regex:k_inflate -c 1 -o $OUT/ncu_inflate_$TAG -f python scripts/probes/inflate_one.py > $OUT/ncu_inflate_$TAG.log 2>&1; tail -2 $OUT/ncu_inflate_$TAG.log
