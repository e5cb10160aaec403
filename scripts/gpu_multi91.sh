N=${1:-8}
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus $N --steps 15 --warmup 3 --read-len 91 > gpurun_out/bench91_n$N.json 2> gpurun_out/bench91_n$N.err
echo "exit $?"; tail -c 300 gpurun_out/bench91_n$N.json
