#!/bin/bash
# tests + kernel-resident bench of the default build and of build/ variants
mkdir -p gpurun_out
TAG=${1:-r2q}; shift
timeout 900 python -m pytest tests -m gpu -q -x -p no:cacheprovider 2>&1 | tail -6
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --parity-reads 200000 > gpurun_out/${TAG}_default.json 2> gpurun_out/${TAG}_default.err; echo "default rc=$?"; tail -2 gpurun_out/${TAG}_default.err
for v in "$@"; do
  PSA_LIB_PATH=$PWD/build/libpsa_$v.so timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --parity-reads 200000 > gpurun_out/${TAG}_$v.json 2> gpurun_out/${TAG}_$v.err; echo "$v rc=$?"
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
