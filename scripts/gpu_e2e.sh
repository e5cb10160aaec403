mkdir -p gpurun_out; : > gpurun_out/e2e.jsonl
for c in 0 131072 262144 524288 2097152; do
  timeout 600 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --chunk-reads $c >> gpurun_out/e2e.jsonl 2>> gpurun_out/e2e.err
done
python - <<'PY'
import json
for l in open('gpurun_out/e2e.jsonl'):
    d=json.loads(l); e=d['e2e']
    print("value %.1fM  e2e %.1fM  %.2f ms/step  clocks %s" % (d['value']/1e6, e['value']/1e6, e['ms_per_step'], d['clocks']))
PY
tail -3 gpurun_out/e2e.err
