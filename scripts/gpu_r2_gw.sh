#!/bin/bash
mkdir -p gpurun_out
TAG=${1:-r2x}
B="--steps 10 --warmup 3 --no-cpu-baseline --no-e2e --parity-reads 200000"
run() { name=$1; shift; env "$@" timeout 300 python bench.py $B $EXTRA > gpurun_out/${TAG}_$name.json 2> gpurun_out/${TAG}_$name.err; echo "$name rc=$?"; tail -1 gpurun_out/${TAG}_$name.err; }
EXTRA="" run base X=0
EXTRA="--group-width 16" run gw16 X=0
EXTRA="--group-width 32" run gw32 X=0
EXTRA="--value-mappers 2" run vm2 X=0
EXTRA="--value-mappers 2 --group-width 16" run vm2_gw16 X=0
EXTRA="--value-mappers 3" run vm3 X=0
python - <<PY
import json,glob
for f in sorted(glob.glob('gpurun_out/${TAG}_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); r=d['roofline']
        print("%-34s %.1fM %.3f ms |"%(f,d['value']/1e6,d['ms_per_step']), " ".join("%s %.3f"%(k,v['ms_per_step']) for k,v in r['kernels'].items()), "| frac %.3f parity"%r['frac'], d.get('parity',{}).get('mismatches'))
    except Exception as e:
        print(f, "ERR", e)
PY
