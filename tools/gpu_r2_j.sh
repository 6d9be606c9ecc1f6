# round-2 call J (1 GPU): bench suite with in-run parity; the reference's own preset shapes (D 768 / 1024, unquantised and FP16 rows)
mkdir -p gpurun_out
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_suite_n1.json 2> gpurun_out/bench_suite_n1.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_suite_n1.err
python - <<'PY'
import json
try:
    d = json.loads(open('gpurun_out/bench_suite_n1.json').read().strip().splitlines()[-1])
    r = d['roofline']
    print('config2', round(d['value']/1e6,1), 'Mtok/s', round(d['ms_per_step']*1e3,2), 'us frac', round(r['frac'],4), 'repeats', [round(x*1e3,2) for x in d['ms_per_step_repeats']], 'isolated', round(r['isolated_kernel_ms']*1e3,2),
          'e2e', round(d['e2e']['value']/1e6,1), d['clocks'], 'parity', d.get('parity'), 'suite_s', d.get('suite_seconds'))
    for k, c in d.get('configs', {}).items():
        if 'error' in c:
            print(k, c)
        else:
            print(k, round(c['value']/1e6,2), 'Mtok/s', round(c['ms_per_step']*1e3,1), 'us frac', round(c['roofline']['frac'],4), c['clocks'], 'wall', round(c['wall_seconds'],1), 'parity', c.get('parity'))
except Exception as e:
    print('parse failed', e)
PY
for spec in custom:768:fp32:3:100000:64:1024:50257 custom:768:fp16:3:100000:64:1024:50257 custom:1024:fp32:3:1000000:64:1024:50257 custom:384:fp16:3:100000:64:1024:50257; do
timeout 300 python tools/tune_embed.py $spec --variants=-1 > gpurun_out/preset_$(echo $spec | tr ':' '_').log 2>&1; grep -E "fused|gather_only|lookup_only|torch_copy_same_bytes\(float32" gpurun_out/preset_$(echo $spec | tr ':' '_').log | cut -c1-200
done
