#!/bin/bash
# round 2, first GPU call: parity tests, a full bench line, kernel-only variants, launch list + full ncu capture
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader; nproc
timeout 900 python -m pytest tests -m gpu -q -x -p no:cacheprovider 2>&1 | tail -15
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/r2a_bench.err
for v in mb4 mb5 mb8 b64mb12; do
  PSA_LIB_PATH=$PWD/build/libpsa_$v.so timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --parity-reads 0 > gpurun_out/r2a_$v.json 2> gpurun_out/r2a_$v.err; echo "$v rc=$?"
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2a_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); r=d['roofline']
        print(f, "reads/s %.1fM ms/step %.3f |"%(d['value']/1e6,d['ms_per_step']), " ".join("%s %.3f"%(k,v['ms_per_step']) for k,v in r['kernels'].items()), "| parity", d.get('parity',{}).get('mismatches'), "e2e", (d.get('e2e') or {}).get('value'))
    except Exception as e:
        print(f, "ERR", e)
PY
ARGS="--steps 2 --warmup 1 --no-cpu-baseline --no-e2e --parity-reads 0 --distinct-batches 1"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_r2a.csv python bench.py $ARGS > gpurun_out/launches_r2a.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_map_lanes|k_seed_scan|k_map' -s 4 -c 4 -f -o gpurun_out/prof_r2a python bench.py $ARGS > gpurun_out/prof_r2a.log 2>&1
ls -la gpurun_out/prof_r2a.ncu-rep; tail -2 gpurun_out/prof_r2a.log
