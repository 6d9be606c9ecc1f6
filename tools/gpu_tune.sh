mkdir -p gpurun_out
for i in 1 2 3; do
for W in config2 config3 config1; do
for PDL in 0 1; do
if [ $PDL = 1 ]; then export SCONE_PDL=1; else unset SCONE_PDL; fi
timeout 200 python tools/tune_embed.py $W --variants=-1 2>&1 | grep -E "fused" | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print('$W PDL=$PDL', round(d['us'],2), round(d['frac'],3))"
done; done; done
