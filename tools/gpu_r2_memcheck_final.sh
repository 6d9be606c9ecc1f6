# round-2 (1 GPU): compute-sanitizer memcheck on the final build -- every embed kernel shape + builders, and the fold tests (all tensor-core paths)
mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool memcheck python tools/sanitize.py > gpurun_out/sanitizer_memcheck_r02.log 2>&1; echo "memcheck rc=$?"; grep -E "ERROR SUMMARY" gpurun_out/sanitizer_memcheck_r02.log | tail -1
timeout 900 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "projection_fold and not 9000 and not 7000 and not 40000" > gpurun_out/sanitizer_memcheck_fold_r02.log 2>&1; echo "memcheck fold rc=$?"; grep -E "passed|failed|ERROR SUMMARY" gpurun_out/sanitizer_memcheck_fold_r02.log | tail -2
