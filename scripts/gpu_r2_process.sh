#!/bin/bash
# round 2: psa_process_reads -- block pipeline (device-side text work) vs host parser; GPU parity tests of both first
mkdir -p gpurun_out
TAG=${1:-r2p}
timeout 900 python -m pytest tests -m gpu -q -x -p no:cacheprovider -k "process_reads" 2>&1 | tail -5
PR_CONFIGS=${PR_CONFIGS:-"PSA_PROCESS_FAST=0;;PSA_FQ_BLOCK_BYTES=16777216;PSA_FQ_LANES=4;PSA_OUT_MMAP=0"} timeout 1500 python scripts/process_reads_bench.py ${N_READS:-8000000} ${PR_THREADS:-16} > gpurun_out/${TAG}_process_reads_full.txt 2>&1
grep -v "^psa:" gpurun_out/${TAG}_process_reads_full.txt | tee gpurun_out/${TAG}_process_reads.txt
grep "^psa:" gpurun_out/${TAG}_process_reads_full.txt | head -${VERBOSE_LINES:-40}
