# session-2 verification: smoke, GPU tests, bench (config 2 + reference arm + config 3), mode timings
mkdir -p gpurun_out
python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke.log
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
timeout 600 python bench.py > gpurun_out/bench_config2.json 2> gpurun_out/bench_config2.err; echo "bench rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/bench_config2.json')); print(d['value']/1e6, d['ms_per_step'], d['roofline']['frac'], d['e2e']['value']/1e6, d['cpu_baseline']['value'], d['gpu_launches'], d['clocks'])"; tail -3 gpurun_out/bench_config2.err
timeout 300 python bench.py --impl reference --steps 10 --warmup 2 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; echo "ref rc=$?"; cut -c1-200 gpurun_out/bench_reference.json
timeout 600 python bench.py --workload config3 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_config3.json 2> gpurun_out/bench_config3.err; python -c "
import json; d=json.load(open('gpurun_out/bench_config3.json')); print('config3', d['value']/1e6, d['ms_per_step'], d['roofline']['frac'], d['e2e']['value']/1e6)"
timeout 300 python tools/tune_embed.py config2 --variants=-1 > gpurun_out/modes_config2.log 2>&1; grep -E "fused|gather_only" gpurun_out/modes_config2.log | cut -c1-200
timeout 300 python tools/tune_embed.py config3 --variants=-1 > gpurun_out/modes_config3.log 2>&1; grep -E "fused|gather_only" gpurun_out/modes_config3.log | cut -c1-200
