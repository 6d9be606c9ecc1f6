mkdir -p gpurun_out
free -g | head -2
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --workload config4 --steps 10 --warmup 3 --rows-per-gpu 2000000 > gpurun_out/bench_config4_w1.json 2> gpurun_out/bench_config4_w1.err; echo "config4(w=1) rc=$?"; cut -c1-1800 gpurun_out/bench_config4_w1.json; tail -3 gpurun_out/bench_config4_w1.err
timeout 900 python bench.py --workload config5 --steps 10 --warmup 3 > gpurun_out/bench_config5.json 2> gpurun_out/bench_config5.err; echo "config5 rc=$?"; cut -c1-1800 gpurun_out/bench_config5.json; tail -3 gpurun_out/bench_config5.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches_bench_config2_r01.csv python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1; tail -1 gpurun_out/bench_under_ncu.log | cut -c1-200
