#!/bin/bash
# A/B of the current build against variants under build/: GPU tests, then kernel-resident bench lines
mkdir -p gpurun_out
TAG=${1:-r2y}; shift
timeout 900 python -m pytest tests -m gpu -q -x -p no:cacheprovider 2>&1 | tail -4
B="--steps 10 --warmup 3 --no-cpu-baseline --no-e2e --parity-reads 1000000"
run() { name=$1; shift; env "$@" timeout 300 python bench.py $B > gpurun_out/${TAG}_$name.json 2> gpurun_out/${TAG}_$name.err; echo "$name rc=$?"; tail -1 gpurun_out/${TAG}_$name.err; }
run head X=0
for v in "$@"; do run $v PSA_LIB_PATH=$PWD/build/libpsa_$v.so; done
run head_b X=0
python - <<PY
import json,glob
for f in sorted(glob.glob('gpurun_out/${TAG}_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); r=d['roofline']
        print("%-34s %.1fM %.3f ms |"%(f,d['value']/1e6,d['ms_per_step']), " ".join("%s %.3f"%(k,v['ms_per_step']) for k,v in r['kernels'].items()), "| frac %.3f parity"%r['frac'], d.get('parity',{}).get('mismatches'))
    except Exception as e:
        print(f, "ERR", e)
PY
