mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 50 --warmup 5 > gpurun_out/bench_config2.json 2> gpurun_out/bench_config2.err; echo "bench rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/bench_config2.json')); print(d['value']/1e6, d['ms_per_step'], d['roofline']['frac'], d['e2e']['mode'], d['e2e']['pipelined']/1e6, d['e2e']['synchronous']/1e6)"; tail -3 gpurun_out/bench_config2.err
timeout 600 python bench.py --steps 200 --warmup 10 --no-cpu-baseline > gpurun_out/bench_config2_k200.json 2> gpurun_out/bench_config2_k200.err; python -c "
import json; d=json.load(open('gpurun_out/bench_config2_k200.json')); print('K=200', d['value']/1e6, d['ms_per_step'], d['roofline']['frac'], d['e2e']['pipelined']/1e6, d['e2e']['synchronous']/1e6)"
