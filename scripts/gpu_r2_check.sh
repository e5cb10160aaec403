#!/bin/bash
# full GPU test suite + the default bench line + the reference arm (what the driver runs at round end)
mkdir -p gpurun_out
TAG=${1:-r2chk}
timeout 1500 python -m pytest tests -m gpu -q -x -p no:cacheprovider 2>&1 | tail -5
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/${TAG}_bench.err
python - <<PY
import json
d=json.loads(open('gpurun_out/${TAG}_bench.json').read().strip().splitlines()[-1]); r=d['roofline']
print("reads/s %.1fM ms/step %.3f |"%(d['value']/1e6,d['ms_per_step']), " ".join("%s %.3f"%(k,v['ms_per_step']) for k,v in r['kernels'].items()), "| frac %.3f traffic/alg %s"%(r['frac'], r.get('traffic_over_algorithmic')))
print("e2e %.1fM packed %.1fM cpu %.2fM"%(d['e2e']['value']/1e6, d['e2e_packed_input']['value']/1e6, d['cpu_baseline']['value']/1e6), "build_s", d['run']['index_build_s'], "setup_s", d['run']['setup_s'])
print({k:d['parity'][k] for k in ('reads_checked','mismatches','checksum_gpu','checksum_cpu')}, d['clocks'])
PY
