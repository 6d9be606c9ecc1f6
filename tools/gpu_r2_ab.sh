# round-2 call AB (1 GPU): fold with two alternating epilogue groups for FP32 / FP16 / INT4 (A/B against SCONE_FOLD_EW=8), parity first
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "projection_fold" 2>&1 | tail -3
timeout 300 python tools/bench_fold.py > gpurun_out/bench_fold_ew16.log 2>&1; echo "ew16 rc=$?"
SCONE_FOLD_EW=8 timeout 300 python tools/bench_fold.py > gpurun_out/bench_fold_ew8.log 2>&1; echo "ew8 rc=$?"
python - <<'PY'
import json
for f in ("ew16", "ew8"):
    for l in open(f"gpurun_out/bench_fold_{f}.log"):
        if l.startswith("{"):
            d = json.loads(l)
            print(f, d["H_f"], d["H"], d["quant"], round(d["ms"], 3), "ms", round(d["TFLOPs_useful"]), "TF", round(d["frac_of_bf16_peak"], 3), "cublas", round(d["cublas_bf16_gemm_only_ms"], 3))
PY
timeout 600 compute-sanitizer --tool racecheck python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "projection_fold and (1280 or 4096)" 2>&1 | tail -4
