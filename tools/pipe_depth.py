"""Development tool: e2e tokens/s of HostPipeline for different slot counts (config 2)."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import scone_b200 as sb  # noqa: E402
from scone_b200.utils import synthetic as S  # noqa: E402

w = bench.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "config2"]
dev = torch.device("cuda", 0)
B, L, D, N, V = w["B"], w["L"], w["D"], w["N"], w["V"]
toks, lens, longest = S.make_vocab_device(N, w["max_n"], V, seed=0, device=dev, return_longest=True)
index = sb.FGramIndex(toks, lens)
table = sb.CacheTable(N, D, w["quant"], device=dev)
S.fill_table_device(table, seed=2)
base = S.make_base_device(V, D, torch.bfloat16, seed=3, device=dev)
hb = [S.make_stream_device(toks, lens, B, L, V, seed=100 + k, p_plant=1.0, pick_ids=longest).cpu().pin_memory() for k in range(8)]
for slots in (2, 3, 4, 6):
    pipe = sb.HostPipeline(index, table, base, (B, L), slots=slots)
    for k in range(10):
        pipe.submit(hb[k % 8])
    pipe.flush()
    torch.cuda.synchronize()
    for rep in range(3):
        t0 = time.perf_counter()
        K = 200
        for k in range(K):
            pipe.submit(hb[k % 8])
        pipe.flush()
        dt = time.perf_counter() - t0
        print(f"slots {slots}: {B * L * K / dt / 1e6:.1f} Mtok/s ({dt / K * 1e6:.1f} us/step)", flush=True)
    pipe.close()
