# round-2 call K (1 GPU): projection fold with CTA pairs that multicast the W tiles (default) against single CTAs (SCONE_FOLD_CLUSTER=1)
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -x -q -k "projection_fold" > gpurun_out/pytest_fold_cluster.log 2>&1; echo "fold pytest (pairs) rc=$?"; tail -3 gpurun_out/pytest_fold_cluster.log | cut -c1-300
SCONE_FOLD_CLUSTER=1 timeout 300 python -m pytest tests -m gpu -x -q -k "projection_fold" > gpurun_out/pytest_fold_single.log 2>&1; echo "fold pytest (single) rc=$?"; tail -3 gpurun_out/pytest_fold_single.log | cut -c1-300
timeout 300 python tools/bench_fold.py > gpurun_out/bench_fold_pairs.log 2>&1; echo "bench_fold pairs rc=$?"
SCONE_FOLD_CLUSTER=1 timeout 300 python tools/bench_fold.py > gpurun_out/bench_fold_single.log 2>&1; echo "bench_fold single rc=$?"
python - <<'PY'
import json
for f in ('gpurun_out/bench_fold_pairs.log', 'gpurun_out/bench_fold_single.log'):
    print(f)
    for l in open(f):
        try:
            d = json.loads(l); print(' ', d['H_f'], d['H'], d['quant'], d['rows'], 'ms', round(d['ms'], 3), 'TF', round(d['TFLOPs_useful']), 'frac', round(d['frac_of_bf16_peak'], 3), 'cublas ms', round(d['cublas_bf16_gemm_only_ms'], 3))
        except Exception:
            print(' ', l.strip()[:200])
PY
