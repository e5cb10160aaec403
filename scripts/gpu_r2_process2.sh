#!/bin/bash
mkdir -p gpurun_out
PR_C3=0 PR_CONFIGS='PSA_VERBOSE=1,REPS=6' timeout 1500 python scripts/process_reads_bench.py ${N_READS:-16000000} 16 > gpurun_out/r2u_full.txt 2>&1
grep -E "summary|^\[|lanes ready" gpurun_out/r2u_full.txt
