#!/bin/bash
# round 2, call M (N GPUs): the driver's multi-GPU launch of both arms + the multi-scene driver's file -> TSDF rates
N=${1:-2}; TAG=${2:-r02m}
OUT=gpurun_out
mkdir -p $OUT
PORT=$((29500 + N))
if [ "$N" = 1 ]; then LAUNCH="python"; else LAUNCH="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $PORT"; fi
( time timeout 1200 $LAUNCH bench.py --gpus $N --steps 20 --warmup 5 > $OUT/bench_n${N}_$TAG.json 2> $OUT/bench_n${N}_$TAG.err ) 2>&1 | tail -3
python - <<PY
import json
j=json.loads(open('$OUT/bench_n${N}_$TAG.json').read().strip().splitlines()[-1])
print('N=$N value', round(j['value']), 'e2e', round(j['e2e']['value']), 'ms', round(j['ms_per_step'],2), 'clocks', j['clocks'])
f=j.get('file_to_tsdf') or {}
print('file_to_tsdf', round(f.get('value',0)), 'many', {k:v for k,v in (f.get('many') or {}).items() if k!='what'})
PY
tail -3 $OUT/bench_n${N}_$TAG.err
if [ "$N" -le 2 ]; then
( time timeout 900 $LAUNCH bench.py --impl reference --gpus $N --steps 3 --warmup 1 > $OUT/bench_ref_n${N}_$TAG.json 2> $OUT/bench_ref_n${N}_$TAG.err ) 2>&1 | tail -3
tail -c 600 $OUT/bench_ref_n${N}_$TAG.json
fi
