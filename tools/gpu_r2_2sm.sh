# round-2 (1 GPU): fold with tcgen05.mma.cta_group::2 as the default for FP32 / FP16 tables -- parity, racecheck, A/B against SCONE_FOLD_2SM=0
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "projection_fold" 2>&1 | tail -4
echo "pytest rc=$?"
timeout 300 python tools/bench_fold.py > gpurun_out/bench_fold_2sm.log 2>&1; echo "default rc=$?"
SCONE_FOLD_2SM=0 timeout 300 python tools/bench_fold.py > gpurun_out/bench_fold_no2sm.log 2>&1; echo "2SM=0 rc=$?"
python - <<'PY'
import json
for f in ("2sm", "no2sm"):
    for l in open(f"gpurun_out/bench_fold_{f}.log"):
        if l.startswith("{"):
            d = json.loads(l)
            print(f, d["H_f"], d["H"], d["quant"], round(d["ms"], 3), "ms", round(d["TFLOPs_useful"]), "TF", round(d["frac_of_bf16_peak"], 3), "cublas", round(d["cublas_bf16_gemm_only_ms"], 3))
PY
SCONE_FOLD_2SM=1 timeout 500 compute-sanitizer --tool racecheck python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "projection_fold and (768 or 448)" 2>&1 | tail -4
timeout 300 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "projection_fold and (1024-517 or 320)" 2>&1 | tail -4
