# repeated A/B of library builds: EXP_SEQ="libA.so - libA.so -" ('-' = the default build)
mkdir -p gpurun_out; : > gpurun_out/exp.jsonl
for lib in $EXP_SEQ; do
  [ "$lib" = "-" ] && lib=""
  PSA_LIB_PATH=${lib:+$PWD/$lib} timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e $EXP_ARGS >> gpurun_out/exp.jsonl 2>> gpurun_out/exp.err
  echo "${lib:-default}" >> gpurun_out/exp.names
done
python - <<'PY'
import json
names=[l.strip() for l in open('gpurun_out/exp.names')][-len(open('gpurun_out/exp.jsonl').readlines()):]
for nm,l in zip(names,open('gpurun_out/exp.jsonl')):
    d=json.loads(l); r=d['roofline']
    print("%-24s reads/s %.1fM  ms/step %.2f | " % (nm, d['value']/1e6, d['ms_per_step']) + "  ".join("%s %.2f" % (k, v['ms_per_launch']) for k, v in r['kernels'].items()))
PY
tail -3 gpurun_out/exp.err; rm -f gpurun_out/exp.names
