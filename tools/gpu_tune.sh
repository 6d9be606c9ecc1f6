mkdir -p gpurun_out
for TOOL in racecheck memcheck; do
timeout 900 compute-sanitizer --tool $TOOL --print-limit 8 python tools/sanitize.py > gpurun_out/sanitizer_$TOOL.log 2>&1; echo "$TOOL rc=$?"; grep -v "^ok" gpurun_out/sanitizer_$TOOL.log | tail -4
done
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -2
for W in config2 config3 config1; do
timeout 200 python tools/tune_embed.py $W --variants=-1 2>&1 | grep -E "fused kind" | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print('$W', round(d['us'],2), round(d['frac'],3))"
done
