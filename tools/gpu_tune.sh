mkdir -p gpurun_out
SCONE_EMBED_VARIANT=1:0:6:12:1:200 timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "embed or config2" > gpurun_out/pytest_bulk.log 2>&1; echo "pytest(bulk 6:12) rc=$?"; tail -3 gpurun_out/pytest_bulk.log
SCONE_EMBED_VARIANT=1:0:2:6:3:70 timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "embed or config2" > gpurun_out/pytest_bulk2.log 2>&1; echo "pytest(bulk 2:6) rc=$?"; tail -3 gpurun_out/pytest_bulk2.log
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
LF=0.25 timeout 200 python tools/tune_embed.py config2 > gpurun_out/tune_config2_lf25.log 2>&1; cat gpurun_out/tune_config2_lf25.log | tail -16
LF=0.25 timeout 300 python tools/tune_embed.py config3 > gpurun_out/tune_config3_lf25.log 2>&1; cat gpurun_out/tune_config3_lf25.log | tail -16
LF=0.25 timeout 200 python tools/tune_embed.py config1 > gpurun_out/tune_config1_lf25.log 2>&1; cat gpurun_out/tune_config1_lf25.log | tail -6
