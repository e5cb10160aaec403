#!/bin/bash
# round 2 evidence run: GPU parity tests, one full bench line, ncu launch list and full capture of the map kernels
mkdir -p gpurun_out
TAG=${1:-r2h}
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader; nproc
timeout 1200 python -m pytest tests -m gpu -q -x -p no:cacheprovider 2>&1 | tail -8
timeout 900 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/${TAG}_bench.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${TAG}_bench_reference.json 2> gpurun_out/${TAG}_bench_reference.err; echo "reference rc=$?"
ARGS="--steps 2 --warmup 1 --no-cpu-baseline --no-e2e --parity-reads 0 --distinct-batches 1"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_$TAG.csv python bench.py $ARGS > gpurun_out/launches_$TAG.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_map|k_seed_scan' -s 5 -c 5 -f -o gpurun_out/prof_$TAG python bench.py $ARGS > gpurun_out/prof_$TAG.log 2>&1
ls -la gpurun_out/prof_$TAG.ncu-rep; tail -2 gpurun_out/prof_$TAG.log
python - <<PY
import json
d=json.loads(open('gpurun_out/${TAG}_bench.json').read().strip().splitlines()[-1]); r=d['roofline']
print("reads/s %.1fM ms/step %.3f |"%(d['value']/1e6,d['ms_per_step']), " ".join("%s %.3f"%(k,v['ms_per_step']) for k,v in r['kernels'].items()), "| frac %.3f"%r['frac'])
print("e2e %.1fM packed %.1fM cpu %s"%(d['e2e']['value']/1e6, d['e2e_packed_input']['value']/1e6, d.get('cpu_baseline')))
print(d['parity'])
PY
