#!/bin/bash
# One gpurun call: parity tests, smoke, small + full bench, machine facts.  Outputs in gpurun_out/.
mkdir -p gpurun_out
{ nproc; free -g | head -2; nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv; } > gpurun_out/machine.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -60 > gpurun_out/pytest_gpu.log
echo "pytest exit: ${PIPESTATUS[0]}" >> gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit: $?" >> gpurun_out/smoke.log
timeout 600 python bench.py --workload gencode_small --steps 5 --warmup 3 --reads-per-step 1048576 > gpurun_out/bench_small.json 2> gpurun_out/bench_small.err
timeout 1200 python bench.py --steps 12 --warmup 3 > gpurun_out/bench_synth.json 2> gpurun_out/bench_synth.err
tail -5 gpurun_out/pytest_gpu.log; cat gpurun_out/smoke.log | tail -3; cat gpurun_out/bench_small.json; tail -3 gpurun_out/bench_small.err; cat gpurun_out/bench_synth.json; tail -3 gpurun_out/bench_synth.err
