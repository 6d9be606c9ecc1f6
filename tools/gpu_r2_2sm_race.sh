mkdir -p gpurun_out
SCONE_FOLD_2SM=1 timeout 500 compute-sanitizer --tool racecheck --print-limit 12 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "projection_fold and 448" > gpurun_out/racecheck_2sm.log 2>&1
grep -E "Race reported|hazard|between|Write|Read|RACECHECK SUMMARY" gpurun_out/racecheck_2sm.log | head -40
