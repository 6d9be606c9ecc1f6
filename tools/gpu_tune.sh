mkdir -p gpurun_out
for TOOL in memcheck racecheck synccheck; do
timeout 900 compute-sanitizer --tool $TOOL --print-limit 5 python tools/sanitize.py > gpurun_out/sanitizer_$TOOL.log 2>&1; echo "$TOOL rc=$?"; tail -3 gpurun_out/sanitizer_$TOOL.log
done
