#!/bin/bash
# round 2 experiment A: (1) L2 fetch granularity probe under ncu, (2) first-pass budgets of the thread kernel
mkdir -p gpurun_out
TAG=${1:-r2i}
M="dram__bytes_read.sum,lts__t_sectors_srcunit_tex_op_read.sum,lts__t_requests_srcunit_tex_op_read.sum,gpu__time_duration.sum"
for g in "" 32 128; do
  ./build/l2_fetch_probe $g > gpurun_out/${TAG}_l2probe_g${g:-def}.txt 2>&1
  timeout 300 ncu --metrics $M --csv --log-file gpurun_out/${TAG}_l2probe_g${g:-def}.csv ./build/l2_fetch_probe $g > /dev/null 2>&1
done
cat gpurun_out/${TAG}_l2probe_gdef.txt
python - <<PY
import csv
for g in ["def","32","128"]:
    rows=list(csv.reader(open("gpurun_out/${TAG}_l2probe_g%s.csv"%g)))
    h=next(i for i,r in enumerate(rows) if r and r[0]=="ID")
    ki,mi,vi=rows[h].index("Kernel Name"),rows[h].index("Metric Name"),rows[h].index("Metric Value")
    d={}
    for r in rows[h+2:]:
        if len(r)>vi: d.setdefault((r[0],r[ki][:30]),{})[r[mi]]=float(r[vi].replace(",",""))
    loads=148*32*256*8
    for (i,k),m in d.items():
        if int(i)%2==1:
            print("gran %-4s id %2s %-28s dram B/load %.1f  L2 sectors/load %.2f  L2 requests/load %.2f  %.3f ms"%(g,i,k,m.get("dram__bytes_read.sum",0)*(1e9 if m.get("dram__bytes_read.sum",0)<1e4 else 1)/loads, m.get("lts__t_sectors_srcunit_tex_op_read.sum",0)/loads, m.get("lts__t_requests_srcunit_tex_op_read.sum",0)/loads, m.get("gpu__time_duration.sum",0)/1e6))
PY
B="--steps 10 --warmup 3 --no-cpu-baseline --no-e2e --parity-reads 200000"
run() { name=$1; shift; env "$@" timeout 300 python bench.py $B $EXTRA > gpurun_out/${TAG}_$name.json 2> gpurun_out/${TAG}_$name.err; echo "$name rc=$?"; }
EXTRA="" run base X=0
for rs in 0 1 2; do EXTRA="" run rs$rs PSA_RESEED_FIRST=$rs; done
for fp in 1 2 4; do EXTRA="--fast-probes $fp" run fp${fp}_rs0 PSA_RESEED_FIRST=0; done
EXTRA="--fast-probes 2" run fp2 X=0
python - <<PY
import json,glob
for f in sorted(glob.glob('gpurun_out/${TAG}_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); r=d['roofline']
        print("%-34s %.1fM %.3f ms |"%(f,d['value']/1e6,d['ms_per_step']), " ".join("%s %.3f"%(k,v['ms_per_step']) for k,v in r['kernels'].items()), "| parity", d.get('parity',{}).get('mismatches'), r.get('handed_over_by_k_map_thread'))
    except Exception as e:
        print(f, "ERR", e)
PY
