mkdir -p gpurun_out; : > gpurun_out/exp.jsonl
for lib in "" gpurun_exp_mb8.so gpurun_exp_mb10.so; do
  for p in 3 8; do
    PSA_LIB_PATH=${lib:+$PWD/$lib} timeout 600 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-e2e --fast-probes $p >> gpurun_out/exp.jsonl 2>> gpurun_out/exp.err
  done
done
python - <<'PY'
import json
for l in open('gpurun_out/exp.jsonl'):
    d=json.loads(l); r=d['roofline']
    print("reads/s %.1fM  ms/step %.2f | " % (d['value']/1e6, d['ms_per_step']) + "  ".join("%s %.2f ms (%d reads)" % (k, v['ms_per_launch'], v['reads_per_launch']) for k, v in r['kernels'].items()))
PY
tail -3 gpurun_out/exp.err
