mkdir -p gpurun_out
for tool in memcheck initcheck; do
  timeout 1200 compute-sanitizer --tool $tool --error-exitcode 9 python -m pytest tests -m gpu -q -x -p no:cacheprovider -k "known_answers or small_fq or tma_read_tiles or wide_classes or device_batch_and_events" > gpurun_out/sanitize_$tool.log 2>&1; echo "$tool exit $?"; grep -E "ERROR SUMMARY|passed|failed" gpurun_out/sanitize_$tool.log | tail -3
done
grep -B2 -A12 "Uninitialized" gpurun_out/sanitize_initcheck.log | head -60
