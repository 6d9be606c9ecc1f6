# round-2 call M: (1 GPU part) fold edge shapes; run with --gpus 4 for the N = 4 suite
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -x -q -k "projection_fold" > gpurun_out/pytest_fold_edges.log 2>&1; echo "fold pytest rc=$?"; tail -4 gpurun_out/pytest_fold_edges.log | cut -c1-300
SCONE_FOLD_CLUSTER=2 timeout 300 python -m pytest tests -m gpu -x -q -k "projection_fold and auto" > gpurun_out/pytest_fold_edges2.log 2>&1; echo "fold pytest (pairs forced) rc=$?"; tail -2 gpurun_out/pytest_fold_edges2.log | cut -c1-300
SCONE_FOLD_CLUSTER=1 timeout 300 python -m pytest tests -m gpu -x -q -k "projection_fold and auto" > gpurun_out/pytest_fold_edges1.log 2>&1; echo "fold pytest (single forced) rc=$?"; tail -2 gpurun_out/pytest_fold_edges1.log | cut -c1-300
bash tools/gpu_r2_f.sh 4
