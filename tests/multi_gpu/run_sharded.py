"""Multi-GPU parity run of the row-sharded tier (NCCL), launched by torchrun on a box with >= 2 GPUs:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29555 tests/multi_gpu/run_sharded.py

Every rank matches its own batch, the table is split by id % W, and the result must equal the C oracle on the unsharded
table, bit for bit.  Not collected by pytest (needs several GPUs); gpurun --gpus 2 runs it.
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

import scone_b200 as sb  # noqa: E402
from scone_b200 import sharded  # noqa: E402
from scone_b200.utils import synthetic as S  # noqa: E402
from oracle import py_oracle as po  # noqa: E402
from oracle.c_oracle import COracleIndex  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
ok = True
for quant, D, max_n in (("int4", 4096, 5), ("fp16", 1024, 3), ("int8", 2048, 4)):
    N, V, B, L = 50_000, 5000, 8, 512
    toks, lens = S.make_vocab_numpy(N, max_n, V, seed=1)
    rows = S.make_rows_numpy(N, D, seed=2)
    base_bits = po.cast_bits(S.make_rows_numpy(V, D, seed=3), "bf16")
    tab = po.OracleTable.from_fp32(rows, quant)
    packed, stride, soff = S.pack_table_numpy(quant, tab.payload, tab.scales)
    q = S.make_stream_numpy(toks, lens, B, L, V, seed=100 + rank)
    want, wid, wlen, err = COracleIndex(toks, lens).embed(quant, D, 128, packed, stride, packed[:, soff:] if soff else None, stride,
                                                          base_bits, q, "bf16", nthreads=4)
    index = sb.FGramIndex(torch.from_numpy(toks).to(dev), torch.from_numpy(lens).to(dev))
    mine = np.arange(rank, N, world)
    table = sb.CacheTable(len(mine), D, quant, device=dev)
    table.store(torch.from_numpy(rows[mine]).to(dev))
    base = torch.from_numpy(base_bits.view(np.int16).copy()).view(torch.bfloat16).to(dev)
    cache = sharded.ShardedEmbeddingCache(sharded.CudaOps(index, table, base), micro_batches=1 if quant == "fp16" else 3)
    emb, fid, ml = cache.lookup(torch.from_numpy(q).to(dev))
    torch.cuda.synchronize()
    got = emb.view(torch.int16).cpu().numpy().view(np.uint16)
    good = np.array_equal(got, want) and np.array_equal(fid.cpu().numpy(), wid) and np.array_equal(ml.cpu().numpy(), wlen)
    print(f"rank {rank}/{world} {quant} D={D}: {'OK' if good else 'MISMATCH'} hits={int((wid >= 0).sum())} micro-batches={cache.last_plan.micro} sent={[sum(c) for c in zip(*cache.last_plan.send_counts)]}", flush=True)
    ok = ok and good
    # peer-direct variant: one kernel, rows pulled over NVLink from symmetric memory
    pt = sharded.PeerShardedTable(N, D, quant, device=dev)
    pt.store_owned(torch.from_numpy(rows[mine]), torch.from_numpy(mine))
    pt.publish()
    status = torch.zeros(1, dtype=torch.int32, device=dev)
    emb2, fid2, ml2 = sharded.embed_forward_sharded(index, pt, base, torch.from_numpy(q).to(dev), status=status)
    torch.cuda.synchronize()
    dist.barrier()
    got2 = emb2.view(torch.int16).cpu().numpy().view(np.uint16)
    good2 = np.array_equal(got2, want) and np.array_equal(fid2.cpu().numpy(), wid) and np.array_equal(ml2.cpu().numpy(), wlen) \
        and int(status.item()) == 0
    print(f"rank {rank}/{world} {quant} D={D} peer-direct: {'OK' if good2 else 'MISMATCH'}", flush=True)
    ok = ok and good2
    # the same with the fused position add: the pipeline kernel pulls the rows from the peers and stages the position rows
    pos_bits = po.cast_bits(S.make_rows_numpy(L, D, seed=7), "bf16")
    want_p, _, _, _ = COracleIndex(toks, lens).embed(quant, D, 128, packed, stride, packed[:, soff:] if soff else None, stride, base_bits, q,
                                                     "bf16", nthreads=4, pos_bits=pos_bits)
    pos = torch.from_numpy(pos_bits.view(np.int16).copy()).view(torch.bfloat16).to(dev)
    emb2p, fid2p, _ = sharded.embed_forward_sharded(index, pt, base, torch.from_numpy(q).to(dev), pos_emb=pos, status=status)
    torch.cuda.synchronize()
    dist.barrier()
    good2p = np.array_equal(emb2p.view(torch.int16).cpu().numpy().view(np.uint16), want_p) and np.array_equal(fid2p.cpu().numpy(), wid) \
        and int(status.item()) == 0
    print(f"rank {rank}/{world} {quant} D={D} peer-direct + wpe: {'OK' if good2p else 'MISMATCH'}", flush=True)
    ok = ok and good2p
# the drop-in class with tier="sharded": every rank offers all rows, keeps its own, looks up through peer memory
ex = sb.NGramExtractor.from_arrays(toks, lens)
cache = sb.EmbeddingCache(ex, D, quant=quant, out_dtype=torch.bfloat16, device=dev, tier="sharded")
cache.cache_embeddings(list(range(N)), torch.from_numpy(rows), verbose=False)
cache.publish()
cache.set_base_embedding(base)
emb3, fid3, _ = cache.lookup(torch.from_numpy(q).to(dev))
torch.cuda.synchronize()
dist.barrier()
good3 = np.array_equal(emb3.view(torch.int16).cpu().numpy().view(np.uint16), want) and np.array_equal(fid3.cpu().numpy(), wid)
print(f"rank {rank}/{world} EmbeddingCache(tier='sharded'): {'OK' if good3 else 'MISMATCH'}", flush=True)
ok = ok and good3
# the reference-API methods take GLOBAL f-gram ids on every rank: rows of other ranks are read through their peer-mapped shard
pick = np.random.default_rng(5 + rank).integers(0, N, size=64)
want_rows = tab.rows_fp32(pick)
got_rows = cache.get_embeddings(pick.tolist()).numpy()
te = cache.get_token_embeddings(q[0, :64].tolist())
g2i = {tuple(int(t) for t in toks[i, :lens[i]]): i for i in range(N)}
tf = po.token_f_grams(g2i.keys(), max_n, q[0, :64].tolist())
good4 = np.array_equal(got_rows, want_rows) and np.array_equal(cache.embeddings[int(pick[0])], want_rows[0]) \
    and all(np.array_equal(te[p_].numpy(), tab.rows_fp32(np.array([g2i[g] for g in gr]))) for p_, gr in tf.items() if gr) \
    and sorted(te) == sorted(p_ for p_, gr in tf.items() if gr)
for what in ("save", "save_binary", "assemble_mean"):
    try:
        getattr(cache, what)("/tmp/should_not_exist") if what != "assemble_mean" else cache.assemble_mean(torch.from_numpy(q).to(dev))
        good4 = False
    except NotImplementedError:
        pass
dist.barrier()
print(f"rank {rank}/{world} sharded get_embeddings / embeddings[] / get_token_embeddings: {'OK' if good4 else 'MISMATCH'}", flush=True)
ok = ok and good4
flag = torch.tensor([0 if ok else 1], device=dev)
dist.all_reduce(flag)
dist.destroy_process_group()
sys.exit(0 if int(flag.item()) == 0 else 1)
