# A/B of alternative builds of the library (PSA_LIB_PATH) on the kernel-resident arm
mkdir -p gpurun_out; : > gpurun_out/exp.jsonl
timeout 1500 python -m pytest tests -m gpu -q -x -p no:cacheprovider 2>&1 | tail -5
for lib in "" $EXP_LIBS; do
  for args in "" $EXP_ARGS; do
    PSA_LIB_PATH=${lib:+$PWD/$lib} timeout 600 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-e2e ${args//,/ } >> gpurun_out/exp.jsonl 2>> gpurun_out/exp.err
  done
done
python - <<'PY'
import json
for l in open('gpurun_out/exp.jsonl'):
    d=json.loads(l); r=d['roofline']
    print("reads/s %.1fM  ms/step %.2f | " % (d['value']/1e6, d['ms_per_step']) + "  ".join("%s %.2f ms (%d reads)" % (k, v['ms_per_launch'], v['reads_per_launch']) for k, v in r['kernels'].items()))
PY
tail -3 gpurun_out/exp.err
