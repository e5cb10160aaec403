#!/bin/bash
# round 2: warp-cooperative first seed search inside k_map_thread (no k_seed_scan for tile batches)
mkdir -p gpurun_out
TAG=${1:-r2v}
timeout 900 python -m pytest tests -m gpu -q -x -p no:cacheprovider 2>&1 | tail -4
B="--steps 10 --warmup 3 --no-cpu-baseline --no-e2e --parity-reads 1000000"
run() { name=$1; shift; env "$@" timeout 300 python bench.py $B $EXTRA > gpurun_out/${TAG}_$name.json 2> gpurun_out/${TAG}_$name.err; echo "$name rc=$?"; tail -1 gpurun_out/${TAG}_$name.err; }
EXTRA="" run coop X=0
EXTRA="" run nocoop PSA_COOP_SEED=0
EXTRA="--fast-probes 1" run coop_fp1 X=0
EXTRA="" run coop_mb12 PSA_LIB_PATH=$PWD/build/libpsa_tmb12.so
EXTRA="--value-mappers 2" run coop_vm2 X=0
EXTRA="" run coop_b X=0
python - <<PY
import json,glob
for f in sorted(glob.glob('gpurun_out/${TAG}_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); r=d['roofline']
        print("%-34s %.1fM %.3f ms |"%(f,d['value']/1e6,d['ms_per_step']), " ".join("%s %.3f"%(k,v['ms_per_step']) for k,v in r['kernels'].items()), "| frac %.3f parity"%r['frac'], d.get('parity',{}).get('mismatches'), r.get('handed_over_by_k_map_thread'))
    except Exception as e:
        print(f, "ERR", e)
PY
