"""Throughput of the projection fold (tcgen05 GEMM + quantising epilogue) against the measured bf16 peak (development tool).

    python tools/bench_fold.py [k_rows]
"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import scone_b200 as sb  # noqa: E402

k = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
dev = torch.device("cuda", 0)
peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
peak = float(peaks.get("bf16_tflops", 1590.0))
hbm = float(peaks.get("hbm_gbs", 6650.0))
two_sweeps = os.environ.get("SCONE_FOLD_XCH") == "0"
for Hf, H, quant in ((384, 768, "fp16"), (384, 1024, "int8"), (768, 1024, "int8"), (768, 1024, "fp16"), (1024, 2048, "int8"), (1024, 4096, "int4"),
                     (1024, 4096, "fp16"), (1024, 4096, "fp32")):
    kk = min(k, (8 << 30) // (4 * H))
    rows = torch.randn((kk, Hf), device=dev, dtype=torch.bfloat16)
    W = torch.randn((H, Hf), device=dev, dtype=torch.bfloat16) * Hf ** -0.5
    t = sb.CacheTable(kk, H, quant, device=dev)
    for _ in range(2):
        t.store_projected(rows, W)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 5
    e0.record()
    for _ in range(reps):
        t.store_projected(rows, W)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    sweeps = 2 if (quant == "int8" and H > 256 and (two_sweeps or H > 2048)) else 1
    flops = 2.0 * kk * Hf * H
    byts = kk * Hf * 2 + kk * t.row_stride
    # what cuBLAS + the separate quantise pass would cost: the library GEMM alone, for reference
    c = torch.empty((kk, H), device=dev, dtype=torch.bfloat16)
    for _ in range(2):
        torch.matmul(rows, W.t(), out=c)
    e0.record()
    for _ in range(reps):
        torch.matmul(rows, W.t(), out=c)
    e1.record()
    torch.cuda.synchronize()
    ms_lib = e0.elapsed_time(e1) / reps
    print(json.dumps({"H_f": Hf, "H": H, "quant": quant, "sweeps": sweeps, "rows": kk, "ms": ms, "TFLOPs_useful": flops / ms / 1e9, "frac_of_bf16_peak": flops / ms / 1e9 / peak,
                      "TFLOPs_issued": sweeps * flops / ms / 1e9, "GBs": byts / ms / 1e6, "frac_of_hbm_peak": byts / ms / 1e6 / hbm,
                      "cublas_bf16_gemm_only_ms": ms_lib, "peak_tflops": peak}), flush=True)
    del rows, W, t, c
