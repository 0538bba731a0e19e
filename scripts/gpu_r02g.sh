#!/bin/bash
# round 2, call G: pipelined integrate A/B, fuse time breakdown against chunk size, ncu source profile of k_inflate
TAG=${1:-r02g}
OUT=gpurun_out
mkdir -p $OUT
for v in cur pipe2_12 pipe2_10 cur pipe2_12 pipe2_10; do
  lib=""; [ $v != cur ] && lib=$PWD/build/ab/$v/libscannet_b200.so
  echo "== $v"; SCN_B200_LIB=$lib timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu --no-seg 2>/dev/null | python -c "import sys,json; j=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(j['value'], j['e2e']['value'], j['ms_per_step'])"
done 2>&1 | tee $OUT/ab_pipe2_$TAG.txt
timeout 600 python scripts/probes/fuse_probe.py 1000 > $OUT/fuse_probe_$TAG.json 2> $OUT/fuse_probe_$TAG.err; python - <<'PY'
import json
j=json.load(open('gpurun_out/fuse_probe_r02g.json'))
for k,v in j.items():
    for r in v: print(k, {x:r[x] for x in ('wall_s','total_s','setup_s','fuse_s','decode_wait_s','depth_decode_s','integrate_s') if x in r})
PY
tail -3 $OUT/fuse_probe_$TAG.err
timeout 900 ncu --set full --import-source on --clock-control none -k regex:k_inflate -c 2 -o $OUT/ncu_inflate_$TAG -f python scripts/probes/inflate_one.py > $OUT/ncu_inflate_$TAG.log 2>&1; tail -3 $OUT/ncu_inflate_$TAG.log
