mkdir -p gpurun_out; : > gpurun_out/gamma.jsonl
for g in 1.7 2.5 4.0; do
  timeout 600 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-e2e --gamma $g >> gpurun_out/gamma.jsonl 2>> gpurun_out/gamma.err
done
python - <<'PY'
import json
for l in open('gpurun_out/gamma.jsonl'):
    d=json.loads(l); r=d['roofline']
    print("levels %d reads/s %.1fM  ms/step %.2f | " % (d['config']['index']['mphf_levels'], d['value']/1e6, d['ms_per_step']) + "  ".join("%s %.2f" % (k, v['ms_per_launch']) for k, v in r['kernels'].items()), "lv/read %.2f" % d['events_per_read']['mphf_levels'])
PY
tail -3 gpurun_out/gamma.err
