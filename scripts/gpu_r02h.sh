#!/bin/bash
# round 2, call H: leaner inflate symbol loop (tests, probe, instruction count), fuse pipeline breakdown with / without SM room left for the decoders
TAG=${1:-r02h}
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m pytest tests/test_inflate_gpu.py tests/test_pipeline_gpu.py -q 2>&1 | tail -8 > $OUT/pytest_inflate_$TAG.log; tail -5 $OUT/pytest_inflate_$TAG.log
timeout 600 python scripts/probes/decode_probe.py inflate > $OUT/decode_probe_$TAG.json 2> $OUT/decode_probe_$TAG.err; python - <<PY
import json
j=json.load(open('$OUT/decode_probe_$TAG.json'))
for k,v in j['inflate'].items(): print(k,v)
PY
tail -3 $OUT/decode_probe_$TAG.err
timeout 600 ncu --metrics smsp__inst_executed.sum,gpu__time_duration.sum,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:k_inflate -c 2 python scripts/probes/inflate_one.py 2>&1 | grep -E "k_inflate|inst_executed|duration|issue_active" > $OUT/ncu_inflate_inst_$TAG.txt; cat $OUT/ncu_inflate_inst_$TAG.txt
timeout 900 python scripts/probes/fuse_probe.py 1000 0 > $OUT/fuse_probe_depth_$TAG.json 2> $OUT/fuse_probe_depth_$TAG.err; tail -2 $OUT/fuse_probe_depth_$TAG.err
timeout 1500 python scripts/probes/fuse_probe.py 2048 1 > $OUT/fuse_probe_c3_$TAG.json 2> $OUT/fuse_probe_c3_$TAG.err; tail -2 $OUT/fuse_probe_c3_$TAG.err
python - <<PY
import json
for f in ('depth','c3'):
    j=json.load(open('$OUT/fuse_probe_%s_$TAG.json' % f))
    for k,v in j.items():
        for r in v: print(f, k, r)
PY
