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
    g = sb.embed_gather(t, base, q, fid)
    m = sb.embed_mean_forward(ix, t, q)
    a = ix.match_all(q)
    torch.cuda.synchronize()
    assert torch.equal(out, g)
    print("ok", quant, D, max_n, "slot bytes", ix.slot_bytes, flush=True)
print("sanitizer workload done")
