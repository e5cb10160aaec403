#!/bin/bash
# round 2: the persistent-lane thread kernel (k_map_walk) -- parity tests, then kernel-resident benches of its variants
mkdir -p gpurun_out
TAG=${1:-r2j}; shift
timeout 900 python -m pytest tests -m gpu -q -x -p no:cacheprovider 2>&1 | tail -6
B="--steps 10 --warmup 3 --no-cpu-baseline --no-e2e --parity-reads 200000"
run() { name=$1; shift; env "$@" timeout 300 python bench.py $B $EXTRA > gpurun_out/${TAG}_$name.json 2> gpurun_out/${TAG}_$name.err; echo "$name rc=$?"; tail -1 gpurun_out/${TAG}_$name.err; }
EXTRA="" run walk16 X=0
EXTRA="" run thread PSA_FAST_KERNEL=thread
for v in "$@"; do EXTRA="" run $v PSA_LIB_PATH=$PWD/build/libpsa_$v.so; done
python - <<PY
import json,glob
for f in sorted(glob.glob('gpurun_out/${TAG}_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); r=d['roofline']
        print("%-34s %.1fM %.3f ms |"%(f,d['value']/1e6,d['ms_per_step']), " ".join("%s %.3f"%(k,v['ms_per_step']) for k,v in r['kernels'].items()), "| parity", d.get('parity',{}).get('mismatches'), r.get('handed_over_by_k_map_thread'))
    except Exception as e:
        print(f, "ERR", e)
PY
