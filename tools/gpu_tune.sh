mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool initcheck --print-limit 12 python tools/sanitize.py > gpurun_out/sanitizer_initcheck.log 2>&1; echo "initcheck rc=$?"; grep -E "Uninit|scone|at::|ERROR SUMMARY" gpurun_out/sanitizer_initcheck.log | sort | uniq -c | sort -rn | head -20
