#!/bin/bash
# one full ncu capture of a named kernel (regex) in a 2-step bench run; usage: gpu_r2_prof1.sh TAG REGEX [ENV=VAL ...]
mkdir -p gpurun_out
TAG=$1; RE=$2; shift; shift
ARGS="--steps 2 --warmup 1 --no-cpu-baseline --no-e2e --parity-reads 0 --distinct-batches 1"
env "$@" timeout 900 ncu --set full --clock-control none --import-source on -k regex:"$RE" -s ${SKIP:-2} -c ${COUNT:-1} -f -o gpurun_out/prof_$TAG python bench.py $ARGS > gpurun_out/prof_$TAG.log 2>&1
ls -la gpurun_out/prof_$TAG.ncu-rep; tail -2 gpurun_out/prof_$TAG.log
