#!/bin/bash
# round 2: blocking thread kernel vs lane pools on the new index layout -- tests under both, benches, ncu of the default
mkdir -p gpurun_out
TAG=${1:-r2c}; shift
timeout 900 python -m pytest tests -m gpu -q -x -p no:cacheprovider 2>&1 | tail -6
PSA_FAST_KERNEL=lanes timeout 900 python -m pytest tests -m gpu -q -x -p no:cacheprovider -k "not config2 and not large_batch" 2>&1 | tail -3
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --parity-reads 200000 > gpurun_out/${TAG}_default.json 2> gpurun_out/${TAG}_default.err; echo "default rc=$?"; tail -2 gpurun_out/${TAG}_default.err
PSA_FAST_KERNEL=lanes timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --parity-reads 0 > gpurun_out/${TAG}_lanes.json 2> gpurun_out/${TAG}_lanes.err; echo "lanes rc=$?"
for v in "$@"; do
  PSA_LIB_PATH=$PWD/build/libpsa_$v.so timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --parity-reads 0 > gpurun_out/${TAG}_$v.json 2> gpurun_out/${TAG}_$v.err; echo "$v rc=$?"
done
python - <<PY
import json,glob
for f in sorted(glob.glob('gpurun_out/${TAG}_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); r=d['roofline']
        print(f, "reads/s %.1fM ms/step %.3f |"%(d['value']/1e6,d['ms_per_step']), " ".join("%s %.3f"%(k,v['ms_per_step']) for k,v in r['kernels'].items()), "| parity", d.get('parity',{}).get('mismatches'))
    except Exception as e:
        print(f, "ERR", e)
PY
ARGS="--steps 2 --warmup 1 --no-cpu-baseline --no-e2e --parity-reads 0 --distinct-batches 1"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_$TAG.csv python bench.py $ARGS > gpurun_out/launches_$TAG.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_map_thread|k_seed_scan|k_map' -s 4 -c 4 -f -o gpurun_out/prof_$TAG python bench.py $ARGS > gpurun_out/prof_$TAG.log 2>&1
ls -la gpurun_out/prof_$TAG.ncu-rep; tail -2 gpurun_out/prof_$TAG.log
