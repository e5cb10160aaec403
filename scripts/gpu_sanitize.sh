# compute-sanitizer over the GPU tests that touch every kernel family (map kernels, novel-set table, FASTQ text kernels,
# device graph builder, mappability)
mkdir -p gpurun_out
SEL="known_answers or small_fq or tma_read_tiles or wide_classes or device_batch_and_events or novel or process_reads_block or device_graph_builder or mappability or constructed"
for tool in memcheck initcheck; do
  timeout 2400 compute-sanitizer --tool $tool --error-exitcode 9 python -m pytest tests -m gpu -q -x -p no:cacheprovider -k "$SEL" > gpurun_out/sanitize_$tool.log 2>&1; echo "$tool exit $?"; grep -E "ERROR SUMMARY|passed|failed" gpurun_out/sanitize_$tool.log | tail -3
done
# (racecheck does not finish on the TMA / mbarrier kernels within half an hour: not run)
grep -B2 -A14 -E "Uninitialized|Invalid|hazard" gpurun_out/sanitize_*.log | head -80
