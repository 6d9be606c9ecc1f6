# round-2 call A: smoke, GPU tests, default bench suite (N = 1), early-start A/B and fused-mode timings
mkdir -p gpurun_out
python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_suite_n1.json 2> gpurun_out/bench_suite_n1.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_suite_n1.err
python - <<'PY'
import json
try:
    d = json.loads(open('gpurun_out/bench_suite_n1.json').read().strip().splitlines()[-1])
    r = d['roofline']
    print('config2', round(d['value']/1e6,1), 'Mtok/s', round(d['ms_per_step']*1e3,2), 'us frac', round(r['frac'],4), 'isolated', round(r['isolated_kernel_ms']*1e3,2), 'us', round(r['isolated_frac'],4),
          'e2e', round(d['e2e']['value']/1e6,1), 'sync', round(d['e2e']['synchronous']/1e6,1), 'to_host', round(d['e2e']['embeds_to_host']['value']/1e6,1), d['clocks'], 'suite_s', d.get('suite_seconds'))
    for k, c in d.get('configs', {}).items():
        if 'error' in c:
            print(k, c)
        else:
            print(k, round(c['value']/1e6,2), 'Mtok/s', round(c['ms_per_step']*1e3,1), 'us frac', round(c['roofline']['frac'],4), c['clocks'], 'wall', round(c['wall_seconds'],1), c['config'].get('f_grams'), c.get('staged'))
except Exception as e:
    print('parse failed', e)
PY
timeout 600 python tools/tune_modes.py config2 "replace" "replace;STABLE=1" "pos" "pos;STABLE=1" "add" "add;STABLE=1" "addpos" "addpos;STABLE=1" "replace" "replace;STABLE=1" > gpurun_out/modes_config2.log 2>&1; cat gpurun_out/modes_config2.log | cut -c1-220
timeout 600 python tools/tune_modes.py config3 "replace" "replace;STABLE=1" "pos" "pos;STABLE=1" "add" "addpos" "addpos;STABLE=1" > gpurun_out/modes_config3.log 2>&1; cat gpurun_out/modes_config3.log | cut -c1-220
timeout 300 python tools/tune_modes.py config1 "replace" "replace;STABLE=1" > gpurun_out/modes_config1.log 2>&1; cat gpurun_out/modes_config1.log | cut -c1-220
