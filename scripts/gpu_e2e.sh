mkdir -p gpurun_out; : > gpurun_out/e2e.jsonl
timeout 1500 python -m pytest tests -m gpu -q -x -p no:cacheprovider 2>&1 | tail -3
for c in ${E2E_CHUNKS:-0 131072 262144 1048576}; do
  timeout 600 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --chunk-reads $c $E2E_ARGS >> gpurun_out/e2e.jsonl 2>> gpurun_out/e2e.err
done
python - <<'PY'
import json
for l in open('gpurun_out/e2e.jsonl'):
    d=json.loads(l); e=d['e2e']
    print("value %.1fM  e2e %.1fM  %.2f ms/step" % (d['value']/1e6, e['value']/1e6, e['ms_per_step']))
PY
tail -3 gpurun_out/e2e.err
