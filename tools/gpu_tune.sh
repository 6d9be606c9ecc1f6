# scratch script of the last experiment run (rewritten per experiment; see profiles/tune_r01.md for the results kept)
mkdir -p gpurun_out
timeout 300 python tools/tune_modes.py config2 "replace;" "pos;" "add;" "addpos;" > gpurun_out/tune_modes_config2.log 2>&1; cut -c1-200 gpurun_out/tune_modes_config2.log | tail -5
