"""Workload for compute-sanitizer (memcheck / racecheck / synccheck): every kernel shape on small inputs."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import scone_b200 as sb  # noqa: E402
from scone_b200.utils import synthetic as S  # noqa: E402

cases = [("int8", 1024, 4, 300), ("int4", 512, 5, 70000), ("fp16", 128, 3, 300), ("int4", 4096, 5, 70000), ("fp16", 8192, 2, 300),
         ("int8", 16384, 3, 300)]
for quant, D, max_n, V in cases:
    toks, lens = S.make_vocab_numpy(1500, max_n, V, seed=1, min_n=1 if max_n < 3 else 2)
    ix = sb.FGramIndex(torch.from_numpy(toks).cuda(), torch.from_numpy(lens).cuda())
    t = sb.CacheTable(1500, D, quant)
    t.store(torch.from_numpy(S.make_rows_numpy(1500, D)).cuda())
    base = torch.randn(V if V < 1000 else 1000, D, device="cuda").to(torch.bfloat16)
    q = torch.from_numpy(S.make_stream_numpy(toks, lens, 3, 130, base.shape[0])).cuda()
    pos = torch.randn(130, D, device="cuda").to(torch.bfloat16)
    out, fid, ml = sb.embed_forward(ix, t, base, q)
    outp, _, _ = sb.embed_forward(ix, t, base, q, pos_emb=pos)
    outa, _, _ = sb.embed_forward(ix, t, base, q, combine="add")
    outap, _, _ = sb.embed_forward(ix, t, base, q, pos_emb=pos, combine="add")
    g = sb.embed_gather(t, base, q, fid)
    m = sb.embed_mean_forward(ix, t, q)
    a = ix.match_all(q)
    torch.cuda.synchronize()
    assert torch.equal(out, g)
    print("ok", quant, D, max_n, "slot bytes", ix.slot_bytes, flush=True)
# a batch large enough for the Bloom pre-filter and the matcher stagger (>= 16 384 positions, >= 2 tiles per matcher)
toks, lens = S.make_vocab_numpy(20000, 4, 5000, seed=3, min_n=2)
ix = sb.FGramIndex(torch.from_numpy(toks).cuda(), torch.from_numpy(lens).cuda())
assert ix.filter_bytes > 0
t = sb.CacheTable(20000, 256, "int8")
t.store(torch.from_numpy(S.make_rows_numpy(20000, 256)).cuda())
base = torch.randn(5000, 256, device="cuda").to(torch.bfloat16)
q = torch.from_numpy(S.make_stream_numpy(toks, lens, 48, 1024, 5000)).cuda()
out, fid, ml = sb.embed_forward(ix, t, base, q)
fid2, ml2 = ix.lookup(q)
torch.cuda.synchronize()
assert torch.equal(fid, fid2) and torch.equal(ml, ml2) and torch.equal(out, sb.embed_gather(t, base, q, fid))
print("ok large batch (pre-filter + stagger)", flush=True)
print("sanitizer workload done")
