# round-2 call C: fold (8 epilogue warps), full GPU tests (3 index/kernel variants), loader-warp sweep for the fused modes, index A/B on config 3
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "projection_fold" > gpurun_out/pytest_fold.log 2>&1; echo "fold pytest rc=$?"; tail -4 gpurun_out/pytest_fold.log | cut -c1-300
timeout 300 python tools/bench_fold.py > gpurun_out/bench_fold.log 2>&1; echo "bench_fold rc=$?"
python - <<'PY'
import json
for l in open('gpurun_out/bench_fold.log'):
    try:
        d = json.loads(l); print(d['H_f'], d['H'], d['quant'], d['rows'], 'ms', round(d['ms'], 3), 'TF', round(d['TFLOPs_useful']), 'frac', round(d['frac_of_bf16_peak'], 3), 'cublas ms', round(d['cublas_bf16_gemm_only_ms'], 3))
    except Exception:
        print(l.strip()[:200])
PY
timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu.log | cut -c1-300
export SCONE_B200_LIB=$PWD/scone_b200/lib/libscone_b200_tune.so
timeout 600 python tools/tune_modes.py config2 "replace;STABLE=1" "pos;STABLE=1,SCONE_EMBED_PIPE=0" "pos;STABLE=1" "add;STABLE=1,SCONE_EMBED_PIPE=0" "add;STABLE=1" "addpos;STABLE=1,SCONE_EMBED_PIPE=0" "addpos;STABLE=1" \
  "pos;STABLE=1,SCONE_EMBED_VARIANT=2:1:6:4:3:70" "pos;STABLE=1,SCONE_EMBED_VARIANT=2:2:5:4:3:70" "pos;STABLE=1,SCONE_EMBED_VARIANT=2:2:4:4:3:70" "pos;STABLE=1,SCONE_EMBED_VARIANT=2:2:4:8:2:100" "pos;STABLE=1,SCONE_EMBED_VARIANT=2:4:6:8:2:100" \
  "add;STABLE=1,SCONE_EMBED_VARIANT=2:2:5:4:3:70" "add;STABLE=1,SCONE_EMBED_VARIANT=2:2:4:8:2:100" "add;STABLE=1,SCONE_EMBED_VARIANT=2:4:6:8:2:100" "add;STABLE=1,SCONE_EMBED_VARIANT=2:4:6:12:1:200" \
  "addpos;STABLE=1,SCONE_EMBED_VARIANT=2:2:6:12:1:200" "addpos;STABLE=1,SCONE_EMBED_VARIANT=2:4:6:12:1:200" "addpos;STABLE=1,SCONE_EMBED_VARIANT=2:4:8:12:1:200" "addpos;STABLE=1,SCONE_EMBED_VARIANT=2:4:8:16:1:200" "addpos;STABLE=1,SCONE_EMBED_VARIANT=2:8:8:16:1:200" "addpos;STABLE=1,SCONE_EMBED_VARIANT=2:4:4:12:1:200" \
  "replace;STABLE=1,SCONE_EMBED_VARIANT=2:2:5:4:3:70" > gpurun_out/modes3_config2.log 2>&1; cut -c1-200 gpurun_out/modes3_config2.log
timeout 600 python tools/tune_modes.py config3 "replace;STABLE=1" "pos;STABLE=1,SCONE_EMBED_PIPE=0" "pos;STABLE=1" "add;STABLE=1" "addpos;STABLE=1,SCONE_EMBED_PIPE=0" "addpos;STABLE=1" \
  "pos;STABLE=1,SCONE_EMBED_VARIANT=2:2:6:12:1:200" "pos;STABLE=1,SCONE_EMBED_VARIANT=2:4:8:12:1:200" "pos;STABLE=1,SCONE_EMBED_VARIANT=2:4:8:16:1:200" "addpos;STABLE=1,SCONE_EMBED_VARIANT=2:4:8:16:1:200" "addpos;STABLE=1,SCONE_EMBED_VARIANT=2:2:6:12:1:200" "replace;STABLE=1,SCONE_EMBED_VARIANT=2:4:6:12:1:200" > gpurun_out/modes3_config3.log 2>&1; cut -c1-200 gpurun_out/modes3_config3.log
echo "--- config 3 with 32-byte slots and no filter (round-1 index)"
SCONE_INDEX_FORMAT=wide SCONE_INDEX_FILTER=never timeout 300 python tools/tune_modes.py config3 "replace;STABLE=1" "replace" > gpurun_out/modes3_config3_wide.log 2>&1; cut -c1-200 gpurun_out/modes3_config3_wide.log
echo "--- config 3 with compact-20 slots, no filter"
SCONE_INDEX_FILTER=never timeout 300 python tools/tune_modes.py config3 "replace;STABLE=1" "replace" > gpurun_out/modes3_config3_c20.log 2>&1; cut -c1-200 gpurun_out/modes3_config3_c20.log
unset SCONE_B200_LIB
