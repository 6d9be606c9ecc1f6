"""Development tool (SCONE_TUNE build): nanosecond timeline of block 0 and the last block of the fused kernel."""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import scone_b200 as sb  # noqa: E402
from scone_b200 import _lib  # noqa: E402
from scone_b200.utils import synthetic as S  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "config2"
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
oid = torch.empty((B, L), dtype=torch.int32, device=dev)
olen = torch.empty((B, L), dtype=torch.uint8, device=dev)
L_ = _lib.load()
names = ["entry", "ids arrived", "first tile resolved", "first bulk issued", "first rows landed", "first position stored", "gather warp 0 done", "matcher 0 done"]
for rep in range(4):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    sb.embed_forward(index, table, base, batches[rep], out=out, out_id=oid, out_len=olen)
    e1.record()
    torch.cuda.synchronize()
    buf = (C.c_ulonglong * 16)()
    L_.scone_debug_timeline(buf)
    t = list(buf)
    t0 = min(t[0], t[8])
    print(f"rep {rep}: events {e0.elapsed_time(e1) * 1e3:.1f} us")
    for blk, off in (("block 0", 0), ("last block", 8)):
        print("  ", blk, ", ".join(f"{names[k]} +{(t[off + k] - t0) / 1e3:.2f}" for k in range(8)))
