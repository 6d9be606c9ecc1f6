mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "wide or hits" > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_gpu.log
for W in config2 config3; do
timeout 300 python tools/tune_embed.py $W --variants=-1 2>&1 | grep -E "variant" | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print(d['variant'], round(d['us'],2), round(d.get('frac',0),3), round(d.get('GBs',0)))"
done
