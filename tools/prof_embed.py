"""A few back-to-back launches of the fused path for ncu (development tool).

    python tools/prof_embed.py [config] [n_steps] [replace|pos|add|addpos] [stable]
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import scone_b200 as sb  # noqa: E402
from scone_b200.utils import synthetic as S  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "config2"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 6
mode = sys.argv[3] if len(sys.argv) > 3 else "replace"
stable = len(sys.argv) > 4 and sys.argv[4] == "stable"
w = bench.WORKLOADS[name]
dev = torch.device("cuda", 0)
B, L, D, N, V = w["B"], w["L"], w["D"], w["N"], w["V"]
toks, lens, longest = S.make_vocab_device(N, w["max_n"], V, seed=0, device=dev, return_longest=True)
index = sb.FGramIndex(toks, lens)
table = sb.CacheTable(N, D, w["quant"], device=dev)
S.fill_table_device(table, seed=2)
base = S.make_base_device(V, D, torch.bfloat16, seed=3, device=dev)
batches = [S.make_stream_device(toks, lens, B, L, V, seed=100 + k, p_plant=1.0, pick_ids=longest) for k in range(4)]
out = torch.empty((B, L, D), dtype=torch.bfloat16, device=dev)
out_id = torch.empty((B, L), dtype=torch.int32, device=dev)
out_len = torch.empty((B, L), dtype=torch.uint8, device=dev)
pos = S.make_base_device(L, D, torch.bfloat16, seed=5, device=dev) if mode in ("pos", "addpos") else None
torch.cuda.synchronize()
for k in range(steps):
    sb.embed_forward(index, table, base, batches[k % 4], out=out, out_id=out_id, out_len=out_len, pos_emb=pos,
                     combine="add" if mode in ("add", "addpos") else "replace", inputs_stable=stable)
torch.cuda.synchronize()
print("done", name, steps, mode, "stable" if stable else "")
