# A/B over environment settings: ENV_AB="A=1 B=2|A=0" (| separates configurations)
mkdir -p gpurun_out; : > gpurun_out/ab.jsonl
timeout 1500 python -m pytest tests -m gpu -q -x -p no:cacheprovider 2>&1 | tail -4
IFS='|' read -ra CFGS <<< "${ENV_AB:-PSA_REG_READS=1|PSA_REG_READS=0}"
for cfg in "${CFGS[@]}"; do
  env $cfg timeout 600 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-e2e $AB_ARGS >> gpurun_out/ab.jsonl 2>> gpurun_out/ab.err
done
python - <<'PY'
import json
for l in open('gpurun_out/ab.jsonl'):
    d=json.loads(l); r=d['roofline']
    print("reads/s %.1fM  ms/step %.2f | " % (d['value']/1e6, d['ms_per_step']) + "  ".join("%s %.2f" % (k, v['ms_per_launch']) for k, v in r['kernels'].items()))
PY
tail -3 gpurun_out/ab.err
