# full round-end style check on the GPU box: tests, smoke, bench (+ reference arm), ncu evidence
mkdir -p gpurun_out
python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 50 --warmup 5 > gpurun_out/bench_config2.json 2> gpurun_out/bench_config2.err; echo "bench rc=$?"; cat gpurun_out/bench_config2.json | cut -c1-1500; tail -3 gpurun_out/bench_config2.err
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; echo "ref rc=$?"; cat gpurun_out/bench_reference.json | cut -c1-600
timeout 900 python bench.py --workload config3 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_config3.json 2> gpurun_out/bench_config3.err; echo "bench3 rc=$?"; cat gpurun_out/bench_config3.json | cut -c1-1500; tail -3 gpurun_out/bench_config3.err
timeout 600 python bench.py --workload config1 --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/bench_config1.json 2> gpurun_out/bench_config1.err; echo "bench1 rc=$?"; cat gpurun_out/bench_config1.json | cut -c1-800
timeout 600 ncu --set full --clock-control none --import-source on -k regex:embed -s 3 -c 1 -o gpurun_out/prof_embed_config2_r01 -f python tools/prof_embed.py config2 6 > gpurun_out/ncu_full2.log 2>&1; tail -2 gpurun_out/ncu_full2.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:embed -s 3 -c 1 -o gpurun_out/prof_embed_config3_r01 -f python tools/prof_embed.py config3 6 > gpurun_out/ncu_full3.log 2>&1; tail -2 gpurun_out/ncu_full3.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_bench_config2_r01.csv python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1; tail -1 gpurun_out/bench_under_ncu.log | cut -c1-200
