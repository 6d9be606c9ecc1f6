mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
timeout 900 python bench.py --workload config5 --steps 10 --warmup 3 --rows-per-gpu 20000000 > gpurun_out/bench_config5.json 2> gpurun_out/bench_config5.err; echo "config5 rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/bench_config5.json')); print('zero-copy', d['value']/1e6, 'Mtok/s', d['roofline']['host_link_GBps'], 'GB/s | staged', d['staged'])"; tail -3 gpurun_out/bench_config5.err
