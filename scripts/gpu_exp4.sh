# A/B of bench settings in one call: lines of "ENV... | bench args" in EXP4 (newline separated)
mkdir -p gpurun_out; : > gpurun_out/exp4.txt; : > gpurun_out/exp4.jsonl
while IFS='|' read -r envs args; do
  [ -z "$envs$args" ] && continue
  env $envs timeout 600 python bench.py --steps 12 --warmup 3 --no-cpu-baseline --no-e2e $args 2>> gpurun_out/exp4.err | tee -a gpurun_out/exp4.jsonl | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('%-40s %-22s reads/s %.1fM  ms/step %.3f | '%('''$envs''','''$args''',d['value']/1e6,d['ms_per_step'])+'  '.join('%s %.3f'%(k,v['ms_per_launch']) for k,v in r['kernels'].items()))" | tee -a gpurun_out/exp4.txt
done <<< "$EXP4"
tail -5 gpurun_out/exp4.err
