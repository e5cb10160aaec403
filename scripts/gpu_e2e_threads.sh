mkdir -p gpurun_out; : > gpurun_out/e2et.jsonl
for t in 1 2 3; do
  timeout 600 python bench.py --steps 12 --warmup 3 --no-cpu-baseline --e2e-threads $t >> gpurun_out/e2et.jsonl 2>> gpurun_out/e2et.err
done
python - <<'PY'
import json
for l in open('gpurun_out/e2et.jsonl'):
    d=json.loads(l); e=d['e2e']
    print("threads %d value %.1fM  e2e %.1fM  %.2f ms/step" % (e['host_threads'], d['value']/1e6, e['value']/1e6, e['ms_per_step']))
PY
tail -3 gpurun_out/e2et.err
