mkdir -p gpurun_out
# reference arm as the driver launches it
timeout 600 python bench.py --impl reference --gpus 1 --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref exit $?"; cut -c1-600 gpurun_out/bench_ref.json
# config 5 shape: 91 bp reads
timeout 600 python bench.py --read-len 91 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_91.json 2> gpurun_out/bench_91.err; echo "91 exit $?"
# memcheck on a subset of the parity tests
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests -m gpu -q -x -p no:cacheprovider -k "known_answers or small_fq or tma_read_tiles or wide_classes or process_reads" > gpurun_out/memcheck.log 2>&1; echo "memcheck exit $?"; tail -5 gpurun_out/memcheck.log
