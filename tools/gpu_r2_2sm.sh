# round-2 experiment (1 GPU): fold with tcgen05.mma.cta_group::2 (SCONE_FOLD_2SM=1), every step under a short timeout
mkdir -p gpurun_out
export SCONE_FOLD_2SM=1
timeout 90 python - <<'PY' 2>&1 | tail -12
import torch, sys
sys.path.insert(0, '.')
import scone_b200 as sb
torch.manual_seed(0)
for (k, Hf, H) in ((256, 64, 256), (300, 64, 256), (1000, 128, 512), (5000, 384, 768)):
    rows = torch.randn(k, Hf, device='cuda'); W = torch.randn(H, Hf, device='cuda') * Hf ** -0.5
    t = sb.CacheTable(k, H, 'fp32')
    t.store_projected(rows, W)
    torch.cuda.synchronize()
    got = t.gather(torch.arange(k, device='cuda'))
    ref = rows.bfloat16().float() @ W.bfloat16().float().t()
    err = (got - ref).abs().max().item()
    print('2sm', k, Hf, H, 'max abs err', err, 'ref absmax', ref.abs().max().item(), flush=True)
PY
echo "direct rc=$?"
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "projection_fold" 2>&1 | tail -4
echo "pytest rc=$?"
timeout 200 python tools/bench_fold.py 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('2sm', d['H_f'], d['H'], d['quant'], round(d['ms'], 3), round(d['TFLOPs_useful']), round(d['frac_of_bf16_peak'], 3), 'cublas', round(d['cublas_bf16_gemm_only_ms'], 3))
"
nvidia-smi --query-gpu=name,memory.used --format=csv,noheader
