# round-2 last call (1 GPU): what the driver runs at round end, on the final tree
mkdir -p gpurun_out
python __graft_entry__.py --smoke 2>&1 | tail -1
timeout 1800 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | cut -c1-300
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_suite_n1_final.json 2> gpurun_out/bench_suite_n1_final.err; echo "bench rc=$?"
python - <<'PY'
import json
d = json.loads(open('gpurun_out/bench_suite_n1_final.json').read().strip().splitlines()[-1])
r = d['roofline']
print('config2', round(d['value']/1e6,1), 'Mtok/s', round(d['ms_per_step']*1e3,2), 'us frac', round(r['frac'],4), 'e2e', round(d['e2e']['value']/1e6,1), d['clocks'], 'parity', d.get('parity', {}).get('result'), 'suite_s', round(d.get('suite_seconds', 0), 1))
for k, c in d.get('configs', {}).items():
    print(k, c.get('error') or (round(c['value']/1e6,2), round(c['ms_per_step']*1e3,1), round(c['roofline']['frac'],4), c['clocks']['reasons'], (c.get('parity') or {}).get('result')))
PY
timeout 300 python tools/bench_fold.py 262144 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print(d['H_f'], d['H'], d['quant'], round(d['ms'], 3), round(d['frac_of_bf16_peak'], 3))
"
