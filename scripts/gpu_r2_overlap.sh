#!/bin/bash
# round 2: k_map's first launch on a second stream (overlap with k_seed_scan + second pass); first-pass re-seed budget
mkdir -p gpurun_out
TAG=${1:-r2l}
timeout 900 python -m pytest tests -m gpu -q -x -p no:cacheprovider 2>&1 | tail -4
B="--steps 10 --warmup 3 --no-cpu-baseline --no-e2e --parity-reads 200000"
run() { name=$1; shift; env "$@" timeout 300 python bench.py $B > gpurun_out/${TAG}_$name.json 2> gpurun_out/${TAG}_$name.err; echo "$name rc=$?"; tail -1 gpurun_out/${TAG}_$name.err; }
run ov1 X=0
run ov0 PSA_OVERLAP_COOP=0
run ov1_rs2 PSA_RESEED_FIRST=2
run ov1_rs3 PSA_RESEED_FIRST=3
run ov1b X=0
run ov0b PSA_OVERLAP_COOP=0
run ov1_rs2b PSA_RESEED_FIRST=2
python - <<PY
import json,glob
for f in sorted(glob.glob('gpurun_out/${TAG}_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); r=d['roofline']
        print("%-34s %.1fM %.3f ms |"%(f,d['value']/1e6,d['ms_per_step']), " ".join("%s %.3f"%(k,v['ms_per_step']) for k,v in r['kernels'].items()), "| parity", d.get('parity',{}).get('mismatches'), r.get('handed_over_by_k_map_thread'))
    except Exception as e:
        print(f, "ERR", e)
PY
