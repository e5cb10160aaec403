#!/bin/bash
# ncu launch list + one full capture of each map kernel (never a bench value: runs under the profiler)
mkdir -p gpurun_out
TAG=${1:-r1}
ARGS="--steps 2 --warmup 1 --no-cpu-baseline --no-e2e --distinct-batches 1 $PROFILE_ARGS"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_$TAG.csv python bench.py $ARGS > gpurun_out/launches_$TAG.log 2>&1
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:'k_map|k_seed_scan' -s 4 -c 4 -f -o gpurun_out/prof_$TAG python bench.py $ARGS > gpurun_out/prof_$TAG.log 2>&1
ls -la gpurun_out/ | tail -4; tail -3 gpurun_out/prof_$TAG.log
