#!/bin/bash
# tests + kernel-resident bench (default + variants) + ncu launch list of the default
mkdir -p gpurun_out
TAG=${1:-r2f}; shift
timeout 900 python -m pytest tests -m gpu -q -x -p no:cacheprovider 2>&1 | tail -4
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --parity-reads 200000 > gpurun_out/${TAG}_default.json 2> gpurun_out/${TAG}_default.err; echo "default rc=$?"; tail -2 gpurun_out/${TAG}_default.err
for v in "$@"; do
  PSA_LIB_PATH=$PWD/build/libpsa_$v.so timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --parity-reads 200000 > gpurun_out/${TAG}_$v.json 2> gpurun_out/${TAG}_$v.err; echo "$v rc=$?"; tail -1 gpurun_out/${TAG}_$v.err
done
python - <<PY
import json,glob
for f in sorted(glob.glob('gpurun_out/${TAG}_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); r=d['roofline']
        print(f, "reads/s %.1fM ms/step %.3f |"%(d['value']/1e6,d['ms_per_step']), " ".join("%s %.3f"%(k,v['ms_per_step']) for k,v in r['kernels'].items()), "| parity", d.get('parity',{}).get('mismatches'), r['handed_over_by_k_map_thread'])
    except Exception as e:
        print(f, "ERR", e)
PY
ARGS="--steps 2 --warmup 1 --no-cpu-baseline --no-e2e --parity-reads 0 --distinct-batches 1"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_$TAG.csv python bench.py $ARGS > gpurun_out/launches_$TAG.log 2>&1
python - <<PY
import csv, collections
rows=list(csv.reader(open('gpurun_out/launches_$TAG.csv')))
h=next(i for i,x in enumerate(rows) if x and x[0]=="ID"); ki,vi=rows[h].index("Kernel Name"),rows[h].index("Metric Value")
tot=collections.defaultdict(float); cnt=collections.defaultdict(int)
for x in rows[h+2:]:
    if len(x)>vi:
        n=x[ki].split("(")[0][:60]; tot[n]+=float(x[vi].replace(",","")); cnt[n]+=1
for k,v in sorted(tot.items(), key=lambda z:-z[1])[:26]:
    print("  %-62s n=%3d avg %.3f ms" % (k, cnt[k], v/cnt[k]/1e6))
PY
