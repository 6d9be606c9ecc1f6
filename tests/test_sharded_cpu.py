"""World-size-2 gloo test (CPU) of the row-sharded tier's host logic: bucketing by owner, the two all-to-alls,
request-order replies, slot map.  The three compute steps are stood in for by the ORACLE (this is a test: the product's
CudaOps has no CPU implementation), so the result must equal the single-process oracle on the unsharded table."""

import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


class OracleOps:
    def __init__(self, toks, lens, packed_local, fmt, base_bits, out_dtype):
        from oracle.c_oracle import COracleIndex
        self.cix = COracleIndex(toks, lens)
        self.packed_local, self.fmt, self.base_bits, self.out_dtype = packed_local, fmt, base_bits, out_dtype

    def match(self, input_ids):
        fid, ml = self.cix.match(input_ids.numpy())
        return torch.from_numpy(fid), torch.from_numpy(ml)

    def serve(self, local_rows):
        return torch.from_numpy(self.packed_local[local_rows.numpy().astype(np.int64)])

    def assemble(self, input_ids, reply, slot_of_position, out):
        from oracle import py_oracle as po
        quant, D, soff = self.fmt
        rb = reply.numpy()
        ids = input_ids.numpy().reshape(-1)
        slot = slot_of_position.numpy()
        res = np.empty((ids.size, D), dtype=np.uint16)
        hit = slot >= 0
        if quant == "int8":
            tab = po.OracleTable("int8", D, rb[:, :D].view(np.int8), rb[:, soff:soff + 4].copy().view(np.float32).reshape(-1))
        elif quant == "int4":
            tab = po.OracleTable("int4", D, rb[:, :D // 2], rb[:, soff:soff + 2 * (D // 128)].copy().view(np.float16))
        else:
            tab = po.OracleTable("fp16", D, rb[:, :2 * D].copy().view(np.float16))
        if hit.any():
            res[hit] = po.cast_bits(tab.rows_fp32(slot[hit]), self.out_dtype)
        res[~hit] = self.base_bits[ids[~hit]]
        res_t = torch.from_numpy(res.reshape(tuple(input_ids.shape) + (D,)).view(np.int16))
        if out is not None:                      # like CudaOps.assemble: the caller's buffer is written in place
            out.copy_(res_t)
            return out
        return res_t


def _worker(rank, world, port, quant, ret):
    import sys
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import py_oracle as po
        from oracle.c_oracle import COracleIndex
        from scone_b200 import sharded
        from scone_b200.utils import synthetic as S
        N, D, V, max_n, B, L = 3001, 128, 200, 4, 3, 77
        toks, lens = S.make_vocab_numpy(N, max_n, V, seed=5)
        rows = S.make_rows_numpy(N, D, seed=6)
        base_bits = po.cast_bits(S.make_rows_numpy(V, D, seed=7), "bf16")
        tab = po.OracleTable.from_fp32(rows, quant)
        packed, stride, soff = S.pack_table_numpy(quant, tab.payload, tab.scales)
        assert sharded.shard_rows(N, rank, world) == len(range(rank, N, world))
        local = packed[rank::world]                                 # rows of the ids this rank owns
        q = S.make_stream_numpy(toks, lens, B, L, V, seed=100 + rank)    # each rank has its own batch
        cache = sharded.ShardedEmbeddingCache(OracleOps(toks, lens, local, (quant, D, soff), base_bits, "bf16"), micro_batches=4)
        emb, fid, ml = cache.lookup(torch.from_numpy(q))
        assert cache.last_plan.micro == 3                              # B = 3 rows: three micro-batches, pipelined
        one = sharded.ShardedEmbeddingCache(cache.ops, micro_batches=1)
        emb1, fid1, _ = one.lookup(torch.from_numpy(q))                 # ... and the unpipelined form gives the same result
        assert one.last_plan.micro == 1 and torch.equal(emb1, emb) and torch.equal(fid1, fid)
        pre = torch.zeros_like(emb)
        emb2, _, _ = cache.lookup(torch.from_numpy(q), out=pre)         # caller-provided output, written per micro-batch
        assert emb2 is pre and torch.equal(pre, emb)
        got = emb.numpy().view(np.uint16)
        want, wid, wlen, err = COracleIndex(toks, lens).embed(quant, D, 128, packed, stride, packed[:, soff:] if soff else None,
                                                              stride, base_bits, q, "bf16")
        ok = err == 0 and np.array_equal(got, want) and np.array_equal(fid.numpy(), wid) and np.array_equal(ml.numpy(), wlen)
        plan = cache.last_plan
        ok = ok and sum(sum(c) for c in plan.send_counts) == int((wid >= 0).sum()) and min(sum(c) for c in zip(*plan.send_counts)) > 0
        # a batch with no hits at all, and one where every hit goes to one owner
        none = torch.full((2, 9), V - 1, dtype=torch.long)
        e2, f2, _ = cache.lookup(none)
        w2, i2, _, _ = COracleIndex(toks, lens).embed(quant, D, 128, packed, stride, packed[:, soff:] if soff else None, stride,
                                                     base_bits, none.numpy(), "bf16")
        ok = ok and np.array_equal(e2.numpy().view(np.uint16), w2) and np.array_equal(f2.numpy(), i2)
        ret[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("quant", ["int8", "int4", "fp16"])
def test_sharded_routing_world2_gloo(quant):
    world = 2
    port = _free_port()
    ctx = mp.get_context("spawn")
    ret = ctx.Manager().dict()
    procs = [ctx.Process(target=_worker, args=(r, world, port, quant, ret)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(180)
        assert p.exitcode == 0
    assert dict(ret) == {0: True, 1: True}


def test_owner_arithmetic():
    from scone_b200 import sharded
    ids = torch.arange(0, 23)
    for W in (1, 2, 8):
        assert torch.equal(sharded.owner_of(ids, W) + W * sharded.local_row_of(ids, W), ids)
        assert sum(sharded.shard_rows(23, r, W) for r in range(W)) == 23
