mkdir -p gpurun_out
rm -rf /dev/shm/psa_gencode_synth_g20000_k24
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ro_base.json 2> gpurun_out/ro.err
rm -rf /dev/shm/psa_gencode_synth_g20000_k24
python scripts/exp_reorder.py 20000 2>> gpurun_out/ro.err
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ro_chain.json 2>> gpurun_out/ro.err
python - <<'PY'
import json
for f in ("ro_base","ro_chain"):
    d=json.load(open("gpurun_out/%s.json"%f)); r=d['roofline']
    print(f, "reads/s %.1fM  ms/step %.2f | " % (d['value']/1e6, d['ms_per_step']) + "  ".join("%s %.2f" % (k, v['ms_per_launch']) for k, v in r['kernels'].items()), d['config']['index']['n_nodes'])
PY
tail -3 gpurun_out/ro.err
