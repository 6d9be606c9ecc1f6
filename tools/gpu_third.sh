set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/pytest_gpu.log
timeout 600 python tools/tune_embed.py config2 --variants=0:0,4:3,4:5,4:6,8:3,8:4,2:6,2:4 > gpurun_out/tune_config2.log 2>&1; cat gpurun_out/tune_config2.log | tail -20
timeout 900 python tools/tune_embed.py config3 --variants=0:0,4:3,4:5,4:6,8:3,8:4,2:6,2:4 > gpurun_out/tune_config3.log 2>&1; cat gpurun_out/tune_config3.log | tail -20
timeout 600 ncu --set full --clock-control none --import-source on -k regex:embed_kernel -s 3 -c 1 -o gpurun_out/prof_embed_config2_v2 -f python tools/prof_embed.py config2 6 > gpurun_out/ncu_full.log 2>&1; tail -3 gpurun_out/ncu_full.log
