#!/bin/bash
# round 2: L2-resident Bloom filter in front of the dictionary in k_seed_scan; with it, shorter first searches in the thread kernel
mkdir -p gpurun_out
TAG=${1:-r2bl}
timeout 900 python -m pytest tests -m gpu -q -x -p no:cacheprovider 2>&1 | tail -4
B="--steps 10 --warmup 3 --no-cpu-baseline --no-e2e --parity-reads 1000000"
run() { name=$1; shift; env "$@" timeout 300 python bench.py $B $EXTRA > gpurun_out/${TAG}_$name.json 2> gpurun_out/${TAG}_$name.err; echo "$name rc=$?"; tail -1 gpurun_out/${TAG}_$name.err; }
EXTRA="" run b10 X=0
for fp in 1 2 3 4 6; do EXTRA="--fast-probes $fp" run b10_fp$fp X=0; done
EXTRA="--fast-probes 2" run b12_fp2 PSA_BLOOM_BITS=12
EXTRA="--fast-probes 2" run b8_fp2 PSA_BLOOM_BITS=8
EXTRA="--fast-probes 2" run b10_fp2_rs1 PSA_RESEED_FIRST=1
EXTRA="--fast-probes 2" run b10_fp2_rs4 PSA_RESEED_FIRST=4
python - <<PY
import json,glob
for f in sorted(glob.glob('gpurun_out/${TAG}_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); r=d['roofline']
        print("%-34s %.1fM %.3f ms |"%(f,d['value']/1e6,d['ms_per_step']), " ".join("%s %.3f"%(k,v['ms_per_step']) for k,v in r['kernels'].items()), "| frac %.3f parity"%r['frac'], d.get('parity',{}).get('mismatches'), r.get('handed_over_by_k_map_thread',{}).get('first_seed_search'))
    except Exception as e:
        print(f, "ERR", e)
PY
