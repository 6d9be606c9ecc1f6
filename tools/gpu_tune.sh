mkdir -p gpurun_out
V="-1,1:0:3:6:3:64,1:0:6:6:2:100,1:0:8:4:2:100,1:0:6:6:3:70,1:0:5:3:4:50,1:0:5:5:3:70,1:0:4:6:3:70"
for SPEC in custom:768:fp16:3:1000000:64:1024:50257 custom:768:int8:3:1000000:64:1024:50257 custom:384:fp16:3:1000000:128:1024:50257; do
echo "== $SPEC"; timeout 200 python tools/tune_embed.py $SPEC --variants=$V 2>&1 | grep -E "load_factor|fused kind|gather_only" | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print('  ', d if 'variant' not in d else (d['variant'], round(d['us'],2), round(d['frac'],3)))"
done
