# psa_process_reads: 8 M reads with per-block timings (PSA_VERBOSE), then 32 M reads (start-up amortised)
mkdir -p gpurun_out
PSA_VERBOSE=1 PR_THREADS=16 bash scripts/gpu_process_reads.sh > gpurun_out/process_reads_verbose.txt 2>&1
grep -v "^psa:" gpurun_out/process_reads_verbose.txt | tail -4
grep "^psa:" gpurun_out/process_reads_verbose.txt | head -24
N_READS=32000000 PR_THREADS=16 bash scripts/gpu_process_reads.sh 2>&1 | tail -5 | tee gpurun_out/process_reads_32m.txt
