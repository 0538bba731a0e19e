#!/bin/bash
# round 2, call T: where the 2.8 ms of the C1 sort go - kernel time per phase (launch-per-phase path under ncu) vs the persistent kernel
TAG=${1:-r02t}
OUT=gpurun_out
mkdir -p $OUT
python scripts/probes/sort_phase_probe.py
SCN_SEG_SORT_LAUNCHES=1 python scripts/probes/sort_phase_probe.py
SCN_SEG_SORT_LAUNCHES=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_ --csv --log-file $OUT/sort_launches_$TAG.csv python scripts/probes/sort_phase_probe.py > $OUT/sort_launches_$TAG.log 2>&1
python - <<PY
import csv, collections
rows=[r for r in csv.reader(open('$OUT/sort_launches_$TAG.csv')) if len(r)>5]
hdr=rows[0]; ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value'); ui=hdr.index('Metric Unit')
acc=collections.Counter(); cnt=collections.Counter()
half=len(rows[1:])//2
for r in rows[1+half:]:       # second (measured) call only
    v=float(r[vi].replace(',','')); v = v/1000 if r[ui]=='ns' else v
    acc[r[ki].split('(')[0]]+=v; cnt[r[ki].split('(')[0]]+=1
for k,v in acc.most_common(): print(f"{k:40s} {cnt[k]:4d} launches {v:9.1f} us")
print('total us', sum(acc.values()))
PY
timeout 600 ncu --set full --clock-control none -k regex:k_sort_levels -c 1 -f -o $OUT/prof_sort_$TAG python scripts/probes/sort_phase_probe.py > $OUT/prof_sort_$TAG.log 2>&1; tail -2 $OUT/prof_sort_$TAG.log
