#!/bin/bash
# N-GPU bench through the driver's own launch line (torchrun), config 4
mkdir -p gpurun_out
N=${1:-2}; TAG=${2:-r2m}
nvidia-smi topo -m > gpurun_out/${TAG}_topo_n$N.txt 2>&1
nproc; free -g | head -2
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/${TAG}_n$N.json 2> gpurun_out/${TAG}_n$N.err
echo "rc=$?"; tail -5 gpurun_out/${TAG}_n$N.err
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/${TAG}_n$N.json').read().strip().splitlines()[-1])
    print("N=%d value %.1fM (%.2f ms/step) e2e %.1fM (%.1f ms) packed %.1fM"%(d['n_gpus'], d['value']/1e6, d['ms_per_step'], d['e2e']['value']/1e6, d['e2e']['ms_per_step'], d['e2e_packed_input']['value']/1e6))
    print(d['parity']); print(d['pcie_probe']); print(d['run']['host_affinity'])
except Exception as e:
    print("ERR", e)
PY
