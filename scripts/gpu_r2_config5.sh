#!/bin/bash
# BASELINE config 5 (scripts/config5.py): N GPUs, reads generated on the device; usage: gpu_r2_config5.sh N READS TAG
mkdir -p gpurun_out
N=${1:-1}; READS=${2:-3e7}; TAG=${3:-r2c5}
if [ "$N" = "1" ]; then
  timeout 1500 python scripts/config5.py --reads $READS > gpurun_out/${TAG}_n$N.json 2> gpurun_out/${TAG}_n$N.err
else
  timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 scripts/config5.py --reads $READS > gpurun_out/${TAG}_n$N.json 2> gpurun_out/${TAG}_n$N.err
fi
echo "rc=$?"; tail -4 gpurun_out/${TAG}_n$N.err; cat gpurun_out/${TAG}_n$N.json
