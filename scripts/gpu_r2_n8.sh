#!/bin/bash
# 8-GPU call: BASELINE config 5 (10^9 reads generated on the device) and the driver-shaped scaling bench at N = 8 (config 4)
mkdir -p gpurun_out
nvidia-smi --query-gpu=name --format=csv,noheader | sort | uniq -c; nproc; free -g | head -2 | tail -1
bash scripts/gpu_r2_config5.sh 8 1e9 r2c5
bash scripts/gpu_r2_multi.sh 8 r2n
