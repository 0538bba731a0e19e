#!/bin/bash
# round 2, call D: knob sweep for the alloc || integrate overlap (SM partitioning), TMA variant
TAG=${1:-r02d}
OUT=gpurun_out
mkdir -p $OUT
B="--steps 10 --warmup 3 --no-cpu --no-seg --parity-frames 0"
run() { name=$1; shift; env "$@" timeout 600 python bench.py $B $EXTRA > $OUT/sweep_${name}_$TAG.json 2> $OUT/sweep_${name}_$TAG.err; python - <<PY
import json
try:
    d=json.load(open("$OUT/sweep_${name}_$TAG.json")); r=d["roofline"]
    print("$name", round(d["value"]), round(d["e2e"]["value"]), "alloc_ms", round(r["alloc_kernel_ms_total"],1), "integ_ms", round(r["integrate_kernel_ms_total"],1), "timed", round(d["timed_region_s"],3))
except Exception as e:
    print("$name", "ERR", open("$OUT/sweep_${name}_$TAG.err").read()[-300:])
PY
}
EXTRA=""
run base X=1
run reserve1 SCN_TSDF_RESERVE=1
run reserve2 SCN_TSDF_RESERVE=2
run reserve3 SCN_TSDF_RESERVE=3
run reserve4 SCN_TSDF_RESERVE=4
run reserve6 SCN_TSDF_RESERVE=6
run group2 SCN_TSDF_ALLOC_GROUP=2
run group8 SCN_TSDF_ALLOC_GROUP=8
run group16 SCN_TSDF_ALLOC_GROUP=16
EXTRA="--tma-kernel"
run tma X=1
run tma_reserve2 SCN_TSDF_RESERVE=2
EXTRA="--batch 32"
run batch32 X=1
run batch32_reserve3 SCN_TSDF_RESERVE=3
EXTRA="--batch 8"
run batch8 X=1
