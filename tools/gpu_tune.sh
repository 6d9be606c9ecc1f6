mkdir -p gpurun_out
for W in config2 config3 config1; do
timeout 200 python tools/tune_embed.py $W --variants=-1 2>&1 | grep -E "fused" | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print('$W', d['variant'], round(d['us'],2), round(d['frac'],3))"
done
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "positions or pipeline or module" 2>&1 | tail -2
