mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
for SPEC in custom:512:int8:4:1000000:128:1024:50257 config1 config2 custom:1024:int4:5:1000000:128:1024:128000 custom:2048:int4:5:2000000:128:1024:128000 custom:2048:int8:5:2000000:128:1024:128000 config3 custom:4096:fp16:5:1000000:64:1024:128000; do
echo "== $SPEC"; LF=0.25 timeout 200 python tools/tune_embed.py $SPEC --variants=-1 2>&1 | grep -E "fused|gather_only" | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print('  ', d['variant'], round(d['us'],2), round(d['frac'],3))"
done
