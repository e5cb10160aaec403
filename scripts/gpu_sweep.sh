#!/bin/bash
# parity tests + a sweep over tuning knobs (kernel-resident arm only)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x -p no:cacheprovider 2>&1 | tail -40 > gpurun_out/pytest_gpu.log
echo "pytest exit: ${PIPESTATUS[0]}" >> gpurun_out/pytest_gpu.log
tail -6 gpurun_out/pytest_gpu.log
: > gpurun_out/sweep.jsonl
SWEEP=${SWEEP:-0:8:16 1:8:16 2:8:16 3:8:16 3:8:0 1:8:32 2:8:32 1:8:8}
for cfg in $SWEEP; do
  IFS=: read p g sw <<< "$cfg"
  timeout 600 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-e2e --fast-probes $p --group-width $g --scan-width ${sw:-16} $SWEEP_ARGS >> gpurun_out/sweep.jsonl 2>> gpurun_out/sweep.err
done
python - <<'PY'
import json
for l in open('gpurun_out/sweep.jsonl'):
    d=json.loads(l); r=d['roofline']
    ks=r.get('kernels',{})
    print("reads/s %.1fM  ms/step %.2f | " % (d['value']/1e6, d['ms_per_step']) +
          "  ".join("%s %.2f ms (%d reads, %.0f GB/s)" % (k, v['ms_per_launch'], v['reads_per_launch'], v['achieved_gbs']) for k, v in ks.items()))
PY
tail -3 gpurun_out/sweep.err
