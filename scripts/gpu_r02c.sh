#!/bin/bash
# round 2, call C: bench A/B (new kernels / r01 kernels / 16 CTAs per SM), parity tests, ncu captures, full bench line, reference arm
TAG=${1:-r02c}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > $OUT/gpu_$TAG.txt 2>&1
nproc >> $OUT/gpu_$TAG.txt; lscpu | grep -E "Model name|Socket|NUMA|Thread" >> $OUT/gpu_$TAG.txt
B="--steps 20 --warmup 5 --no-cpu --no-seg"
( time timeout 900 python bench.py $B > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err ) 2>&1 | tail -3; cat $OUT/bench_$TAG.json; tail -5 $OUT/bench_$TAG.err
SCN_B200_LIB=$PWD/build/ab/r01/libscannet_b200.so timeout 900 python bench.py $B > $OUT/bench_r01kernels_$TAG.json 2> $OUT/bench_r01kernels_$TAG.err; cat $OUT/bench_r01kernels_$TAG.json; tail -3 $OUT/bench_r01kernels_$TAG.err
SCN_B200_LIB=$PWD/build/ab/ctas16/libscannet_b200.so timeout 900 python bench.py $B --parity-frames 0 > $OUT/bench_ctas16_$TAG.json 2> $OUT/bench_ctas16_$TAG.err; cat $OUT/bench_ctas16_$TAG.json; tail -3 $OUT/bench_ctas16_$TAG.err
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -60 > $OUT/pytest_gpu_$TAG.log
cat $OUT/pytest_gpu_$TAG.log
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu --no-seg --color > $OUT/bench_color_$TAG.json 2> $OUT/bench_color_$TAG.err; cat $OUT/bench_color_$TAG.json; tail -3 $OUT/bench_color_$TAG.err
N="--steps 2 --warmup 1 --scene-frames 96 --frames-per-step 96 --no-cpu --no-seg --parity-frames 0"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_ -c 200 --csv --log-file $OUT/launches_$TAG.csv python bench.py $N > $OUT/ncu_launch_$TAG.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_integrate -s 8 -c 2 -f -o $OUT/prof_integrate_$TAG python bench.py $N > $OUT/ncu_full_$TAG.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_alloc -s 8 -c 1 -f -o $OUT/prof_alloc_$TAG python bench.py $N > $OUT/ncu_full_alloc_$TAG.log 2>&1
tail -3 $OUT/ncu_full_$TAG.log
( time timeout 1500 python bench.py --steps 20 --warmup 5 > $OUT/bench_full_$TAG.json 2> $OUT/bench_full_$TAG.err ) 2>&1 | tail -3; cat $OUT/bench_full_$TAG.json; tail -8 $OUT/bench_full_$TAG.err
( time timeout 900 python bench.py --impl reference --steps 20 --warmup 5 > $OUT/bench_ref_$TAG.json 2> $OUT/bench_ref_$TAG.err ) 2>&1 | tail -3; tail -c 1800 $OUT/bench_ref_$TAG.json
