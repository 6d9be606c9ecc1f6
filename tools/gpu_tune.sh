mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
V="-1,1:0:3:6:3:64,1:0:4:8:2:100,1:0:6:6:2:100,1:0:3:5:3:70,1:0:6:10:2:100,1:0:12:12:1:200"
for LF in 0.25 0.4 0.5; do
echo "== config2 LF=$LF"; LF=$LF timeout 200 python tools/tune_embed.py config2 --variants=$V 2>&1 | grep -E "load_factor|lookup_only|fused" | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l)
    print(d if 'variant' not in d else (d['variant'], round(d['us'],2), round(d['frac'],3)))"
done
for LF in 0.25 0.5; do
echo "== config3 LF=$LF"; LF=$LF timeout 300 python tools/tune_embed.py config3 --variants=-1,1:0:6:10:2:100,1:0:12:20:1:200,1:0:8:16:1:200 2>&1 | grep -E "load_factor|lookup_only|fused" | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l)
    print(d if 'variant' not in d else (d['variant'], round(d['us'],2), round(d['frac'],3)))"
done
