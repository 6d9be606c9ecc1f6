mkdir -p gpurun_out
for A in 32 64 128; do for W in config2 config3; do
ALIGN=$A timeout 200 python tools/tune_embed.py $W --variants=-1 2>&1 | grep -E "row_stride|fused kind|gather_only" | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print('$W align=$A', d if 'variant' not in d else (d['variant'][:12], round(d['us'],2), round(d['frac'],3)))"
done; done
