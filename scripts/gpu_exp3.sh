mkdir -p gpurun_out; : > gpurun_out/exp.txt
for lib in $EXP_SEQ; do
  [ "$lib" = "-" ] && lib=""
  for wl in gencode_synth gencode_small; do
    extra=""; [ $wl = gencode_small ] && extra="--reads-per-step 1048576"
    PSA_LIB_PATH=${lib:+$PWD/$lib} timeout 600 python bench.py --workload $wl $extra --steps 10 --warmup 3 --no-cpu-baseline --no-e2e 2>> gpurun_out/exp.err | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('%-22s %-14s reads/s %.1fM  ms/step %.2f | '%('${lib:-default}','$wl',d['value']/1e6,d['ms_per_step'])+'  '.join('%s %.2f'%(k,v['ms_per_launch']) for k,v in r['kernels'].items()))" | tee -a gpurun_out/exp.txt
  done
done
tail -3 gpurun_out/exp.err
