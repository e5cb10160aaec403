# parity tests + one kernel-resident bench + launch list shares
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x -p no:cacheprovider 2>&1 | tail -4
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e $QUICK_ARGS > gpurun_out/quick.json 2> gpurun_out/quick.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_quick.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --distinct-batches 1 > /dev/null 2>&1
python - <<'PY'
import json, csv, collections
d=json.load(open('gpurun_out/quick.json')); r=d['roofline']
print("reads/s %.1fM  ms/step %.2f | " % (d['value']/1e6, d['ms_per_step']) + "  ".join("%s %.2f" % (k, v['ms_per_launch']) for k, v in r['kernels'].items()))
rows=list(csv.reader(open('gpurun_out/launches_quick.csv')))
h=next(i for i,x in enumerate(rows) if x and x[0]=="ID"); ki,vi=rows[h].index("Kernel Name"),rows[h].index("Metric Value")
tot=collections.defaultdict(float); cnt=collections.defaultdict(int)
for x in rows[h+2:]:
    if len(x)>vi:
        n=x[ki].split("(")[0][:44]; tot[n]+=float(x[vi].replace(",","")); cnt[n]+=1
for k,v in sorted(tot.items(), key=lambda z:-z[1])[:12]:
    print("  %-46s n=%3d avg %.3f ms" % (k, cnt[k], v/cnt[k]/1e6))
PY
tail -3 gpurun_out/quick.err
