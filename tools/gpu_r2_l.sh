# round-2 call L (1 GPU): host pipeline on two alternating compute streams
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "pipeline or dropin or input_embedding" > gpurun_out/pytest_pipe.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_pipe.log | cut -c1-200
for i in 1 2; do
timeout 600 python bench.py --workload config2 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_config2_l$i.json 2> gpurun_out/bench_config2_l$i.err; echo "bench rc=$?"
python - $i <<'PY'
import json, sys
d = json.loads(open(f'gpurun_out/bench_config2_l{sys.argv[1]}.json').read().strip().splitlines()[-1])
print('config2', round(d['value']/1e6,1), 'Mtok/s', round(d['ms_per_step']*1e3,2), 'us frac', round(d['roofline']['frac'],4), 'e2e pipelined', round(d['e2e']['pipelined']/1e6,1), 'sync', round(d['e2e']['synchronous']/1e6,1), d['clocks'], d['parity']['result'])
PY
done
timeout 600 python bench.py --workload config3 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_config3_l.json 2> gpurun_out/bench_config3_l.err; python - <<'PY'
import json
d = json.loads(open('gpurun_out/bench_config3_l.json').read().strip().splitlines()[-1])
print('config3', round(d['value']/1e6,1), 'Mtok/s', round(d['ms_per_step']*1e3,1), 'us frac', round(d['roofline']['frac'],4), 'e2e pipelined', round(d['e2e']['pipelined']/1e6,1), 'sync', round(d['e2e']['synchronous']/1e6,1), d['clocks'])
PY
