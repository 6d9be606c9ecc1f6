# round-2 call Y (1 GPU): INT8 projection fold in one sweep (row absmax exchanged inside a cluster): parity first, under a short timeout, then throughput
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "projection_fold" 2>&1 | tail -5
timeout 300 python tools/bench_fold.py > gpurun_out/bench_fold_xch.log 2>&1; echo "fold xch rc=$?"
SCONE_FOLD_XCH=0 timeout 300 python tools/bench_fold.py > gpurun_out/bench_fold_2sweep.log 2>&1; echo "fold 2-sweep rc=$?"
python - <<'PY'
import json
for f in ("xch", "2sweep"):
    for l in open(f"gpurun_out/bench_fold_{f}.log"):
        if l.startswith("{"):
            d = json.loads(l)
            print(f, d["H_f"], d["H"], d["quant"], "sweeps", d["sweeps"], round(d["ms"], 3), "ms", round(d["TFLOPs_useful"]), "TF", round(d["frac_of_bf16_peak"], 3), "cublas", round(d["cublas_bf16_gemm_only_ms"], 3))
PY
timeout 600 compute-sanitizer --tool racecheck python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "projection_fold and 1280" 2>&1 | tail -6
