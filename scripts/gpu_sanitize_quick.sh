mkdir -p gpurun_out
timeout 1500 compute-sanitizer --tool initcheck --error-exitcode 9 python -m pytest tests -m gpu -q -x -p no:cacheprovider -k "novel or constructed or process_reads_block or device_graph_builder or mappability" > gpurun_out/sanitize_initcheck2.log 2>&1; echo "initcheck exit $?"; grep -E "ERROR SUMMARY|passed|failed" gpurun_out/sanitize_initcheck2.log | tail -3
grep "Device Frame\|Host Frame: [a-z_]*(" gpurun_out/sanitize_initcheck2.log | sort | uniq -c | sort -rn | head -8
