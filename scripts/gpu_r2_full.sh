#!/bin/bash
# tests + one FULL bench line (value, e2e arms, cpu baseline, parity) + variants kernel-only
mkdir -p gpurun_out
TAG=${1:-r2e}; shift
timeout 900 python -m pytest tests -m gpu -q -x -p no:cacheprovider 2>&1 | tail -6
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/${TAG}_full.json 2> gpurun_out/${TAG}_full.err; echo "full rc=$?"; tail -3 gpurun_out/${TAG}_full.err
for v in "$@"; do
  PSA_LIB_PATH=$PWD/build/libpsa_$v.so timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --parity-reads 200000 > gpurun_out/${TAG}_$v.json 2> gpurun_out/${TAG}_$v.err; echo "$v rc=$?"; tail -1 gpurun_out/${TAG}_$v.err
done
python - <<PY
import json,glob
for f in sorted(glob.glob('gpurun_out/${TAG}_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); r=d['roofline']
        print(f, "reads/s %.1fM ms/step %.3f |"%(d['value']/1e6,d['ms_per_step']), " ".join("%s %.3f"%(k,v['ms_per_step']) for k,v in r['kernels'].items()), "| parity", d.get('parity',{}).get('mismatches'), d.get('parity',{}).get('novel_sets'))
        if d.get('e2e'): print("   e2e %.1fM (%.1f ms)  packed %.1fM  cpu %.1fM  probe %s"%(d['e2e']['value']/1e6, d['e2e']['ms_per_step'], d['e2e_packed_input']['value']/1e6, (d.get('cpu_baseline') or {}).get('value',0)/1e6, d.get('pcie_probe')))
    except Exception as e:
        print(f, "ERR", e)
PY
