mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
for i in 1 2; do
timeout 300 python bench.py --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/bench_pdl.json 2> gpurun_out/bench_pdl.err; python -c "
import json; d=json.load(open('gpurun_out/bench_pdl.json')); print('PDL   ', d['ms_per_step']*1e3, d['roofline']['frac'], d['e2e']['value']/1e6)"
SCONE_NO_PDL=1 timeout 300 python bench.py --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/bench_nopdl.json 2> gpurun_out/bench_nopdl.err; python -c "
import json; d=json.load(open('gpurun_out/bench_nopdl.json')); print('no PDL', d['ms_per_step']*1e3, d['roofline']['frac'], d['e2e']['value']/1e6)"
done
timeout 300 python bench.py --workload config3 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_config3.json 2>gpurun_out/bench_config3.err; python -c "
import json; d=json.load(open('gpurun_out/bench_config3.json')); print('config3 PDL', d['ms_per_step']*1e3, d['roofline']['frac'], d['e2e']['value']/1e6)"
tail -2 gpurun_out/bench_pdl.err
