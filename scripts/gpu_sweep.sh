#!/bin/bash
# parity tests + a sweep over tuning knobs (kernel-resident arm only)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -30 > gpurun_out/pytest_gpu.log
echo "pytest exit: ${PIPESTATUS[0]}" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
: > gpurun_out/sweep.jsonl
for g in 8 16 32; do
  timeout 600 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-e2e --group-width $g $SWEEP_ARGS >> gpurun_out/sweep.jsonl 2>> gpurun_out/sweep.err
done
python - <<'PY'
import json
for l in open('gpurun_out/sweep.jsonl'):
    d=json.loads(l); r=d['roofline']
    print("reads/s %.1fM  ms/step %.2f  k_map ms %.2f share %.2f" % (d['value']/1e6, d['ms_per_step'], r['kernel_ms_per_launch'], r['kernel_share_of_step']))
PY
tail -3 gpurun_out/sweep.err
