mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
for LF in 0 0.25 0.35; do for FMT in auto wide; do
echo "== config2 LF=$LF fmt=$FMT"; if [ $FMT = wide ]; then export SCONE_INDEX_FORMAT=wide; else unset SCONE_INDEX_FORMAT; fi
LF=$LF timeout 200 python tools/tune_embed.py config2 --variants=-1 2>&1 | grep -E "load_factor|lookup_only|fused|gather_only" | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print('  ', d if 'variant' not in d else (d['variant'], round(d['us'],2), round(d['frac'],3)))"
done; done
unset SCONE_INDEX_FORMAT
echo "== config1"; timeout 200 python tools/tune_embed.py config1 --variants=-1 2>&1 | grep -E "load_factor|lookup_only|fused|gather_only" | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print('  ', d if 'variant' not in d else (d['variant'], round(d['us'],2), round(d['frac'],3)))"
