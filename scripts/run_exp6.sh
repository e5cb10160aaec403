# A/B harness of one experiment batch (results: profiles/r1_exp_overlap_knobs.txt).  The libpsa_*.so variants under build/exp/
# are built first with: make -C rust-pseudoaligner_b200/csrc OUT=../../build/exp/libpsa_<name>.so NVFLAGS="<flags> -D<switch>=<value>" ../../build/exp/libpsa_<name>.so
# end-to-end arm: host threads (one mapper each) and pipeline chunk size
mkdir -p gpurun_out; : > gpurun_out/exp6.txt
for cfg in "2 0" "3 0" "4 0" "2 1048576" "3 262144"; do
  set -- $cfg
  timeout 600 python bench.py --steps 12 --warmup 3 --no-cpu-baseline --e2e-threads $1 --chunk-reads $2 2>> gpurun_out/exp6.err | python -c "
import json,sys
d=json.loads(sys.stdin.read()); e=d['e2e']
print('e2e threads $1 chunk $2: %.1f M reads/s, %.2f ms/step | resident %.1f M' % (e['value']/1e6, e['ms_per_step'], d['value']/1e6))" | tee -a gpurun_out/exp6.txt
done
tail -3 gpurun_out/exp6.err
# occupancy data point for k_map_thread: 72 registers, no spills, 14 CTAs x 64 threads = 896 threads per SM
PSA_LIB_PATH=$PWD/build/exp/libpsa_occ12.so timeout 600 python bench.py --steps 12 --warmup 3 --no-cpu-baseline --no-e2e 2>> gpurun_out/exp6.err | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('occ12 (896 threads/SM, 72 regs): reads/s %.1fM  ms/step %.3f | '%(d['value']/1e6,d['ms_per_step'])+'  '.join('%s %.3f'%(k,v['ms_per_launch']) for k,v in r['kernels'].items()))" | tee -a gpurun_out/exp6.txt
