mkdir -p gpurun_out
: > gpurun_out/diag.jsonl
for cfg in 1:32 3:32 8:32 3:1000; do
  p=${cfg%%:*}; s=${cfg##*:}
  timeout 600 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-e2e --fast-probes $p --fast-max-small $s >> gpurun_out/diag.jsonl 2>> gpurun_out/diag.err
done
python - <<'PY'
import json
for l in open('gpurun_out/diag.jsonl'):
    d=json.loads(l); r=d['roofline']
    print("%.1fM" % (d['value']/1e6), r['handed_over_by_k_map_thread'], {k: round(v['ms_per_launch'],2) for k,v in r['kernels'].items()})
PY
tail -3 gpurun_out/diag.err
