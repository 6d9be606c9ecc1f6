# round-2 call R (1 GPU): new mode defaults -- full GPU tests, then the four modes on configs 1-3 with the library defaults
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log | cut -c1-200
export SCONE_B200_LIB=$PWD/scone_b200/lib/libscone_b200_tune.so
timeout 600 python tools/tune_modes.py config2 "replace;STABLE=1" "pos;STABLE=1" "add;STABLE=1" "addpos;STABLE=1" "pos;STABLE=1,SCONE_EMBED_VARIANT=2:2:4:8:2:110" "pos;STABLE=1,SCONE_EMBED_VARIANT=2:2:6:8:2:110" "replace" "pos" "add" "addpos" > gpurun_out/modes7_config2.log 2>&1; cut -c1-200 gpurun_out/modes7_config2.log
timeout 600 python tools/tune_modes.py config3 "replace;STABLE=1" "pos;STABLE=1" "add;STABLE=1" "addpos;STABLE=1" > gpurun_out/modes7_config3.log 2>&1; cut -c1-200 gpurun_out/modes7_config3.log
timeout 600 python tools/tune_modes.py config1 "replace;STABLE=1" "pos;STABLE=1" "add;STABLE=1" "addpos;STABLE=1" > gpurun_out/modes7_config1.log 2>&1; cut -c1-200 gpurun_out/modes7_config1.log
