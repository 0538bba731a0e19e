#!/bin/bash
# round 2, call P: final kernels of the round (k_alloc diet, lean inflate, MCU IDCT, pooled staging): tests, full bench line, ncu, reference arm
TAG=${1:-r02s}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > $OUT/gpu_$TAG.txt 2>&1
nproc >> $OUT/gpu_$TAG.txt; lscpu | grep -E "Model name|Socket|NUMA|Thread" >> $OUT/gpu_$TAG.txt
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -60 > $OUT/pytest_gpu_$TAG.log
cat $OUT/pytest_gpu_$TAG.log | tail -15
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke_$TAG.log 2>&1; tail -6 $OUT/smoke_$TAG.log
( time timeout 1500 python bench.py --steps 20 --warmup 5 > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err ) 2>&1 | tail -3; cat $OUT/bench_$TAG.json; tail -8 $OUT/bench_$TAG.err
N="--steps 2 --warmup 1 --scene-frames 96 --frames-per-step 96 --no-cpu --no-seg --parity-frames 0"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_ -c 200 --csv --log-file $OUT/launches_$TAG.csv python bench.py $N > $OUT/ncu_launch_$TAG.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_integrate -s 4 -c 2 -f -o $OUT/prof_integrate_$TAG python bench.py $N > $OUT/ncu_full_$TAG.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_alloc -s 4 -c 1 -f -o $OUT/prof_alloc_$TAG python bench.py $N > $OUT/ncu_full_alloc_$TAG.log 2>&1
( time timeout 900 python bench.py --impl reference --steps 20 --warmup 5 > $OUT/bench_ref_$TAG.json 2> $OUT/bench_ref_$TAG.err ) 2>&1 | tail -3; tail -c 1200 $OUT/bench_ref_$TAG.json
SCN_SEG_SORT_LAUNCHES=1 timeout 600 python - > $OUT/sort_ab_$TAG.log 2>&1 <<'PY'
import sys, time
sys.path.insert(0, "."); sys.path.insert(0, "tests")
from scannet_b200 import segmentator, synth
for nx, ny in ((250, 200), (1600, 1250)):
    xyz, tri = synth.make_feature_mesh(nx, ny, seed=5)
    segmentator.segment_mesh(xyz, tri)
    segmentator.segment_mesh(xyz, tri)
    ms, n = segmentator.last_timings()
    print("launch-per-phase sort", nx * ny, "verts: sort", round(ms[3], 3), "ms,", n, "launches")
PY
cat $OUT/sort_ab_$TAG.log
