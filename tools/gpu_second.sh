set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/pytest_gpu.log
timeout 600 python tools/tune_embed.py config2 > gpurun_out/tune_config2.log 2>&1; cat gpurun_out/tune_config2.log | tail -20
timeout 600 ncu --set full --clock-control none --import-source on -k regex:embed_kernel -s 3 -c 2 -o gpurun_out/prof_embed_config2 -f python tools/prof_embed.py config2 6 > gpurun_out/ncu_full.log 2>&1; tail -3 gpurun_out/ncu_full.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_bench_config2.csv python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1; tail -2 gpurun_out/bench_under_ncu.log
timeout 900 python tools/tune_embed.py config3 > gpurun_out/tune_config3.log 2>&1; cat gpurun_out/tune_config3.log | tail -20
