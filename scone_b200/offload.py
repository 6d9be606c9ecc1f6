"""Offloaded tier: cache rows too large for HBM live in pinned host memory.

Two ways to read them (BASELINE.json north_star item 3):

* **zero-copy** -- nothing to do here: a ``CacheTable(tier="host")`` hands the kernels a UVA-mapped pointer and the fused
  kernel's TMA bulk copies pull each matched row over the host link straight into shared memory.
* **staged** (this module) -- the batch is cut into micro-batches; for each one the match runs on the GPU, the matched row
  numbers come back to the host, host threads gather those rows into a pinned staging buffer
  (``scone_host_gather_rows``), one ``cudaMemcpyAsync`` on a side stream moves it to HBM, and ``scone_embed_gather``
  dequantises it.  The host gather + copy of micro-batch m+1 overlap the GPU work of micro-batch m.
"""

from __future__ import annotations

import os
from typing import Optional

import torch

from . import _lib
from .index import FGramIndex
from .table import CacheTable, embed_gather


class StagedHostLookup:
    def __init__(self, index: FGramIndex, host_table: CacheTable, base_emb: torch.Tensor, micro_batches: int = 4,
                 max_positions: int = 1 << 20, threads: Optional[int] = None):
        if host_table.tier != "host":
            raise ValueError("StagedHostLookup needs a tier='host' table")
        self.index, self.table, self.base = index, host_table, base_emb
        self.m = max(1, micro_batches)
        self.threads = threads or (os.cpu_count() or 1)
        dev = index.device
        cap = (max_positions + self.m - 1) // self.m
        stride = host_table.row_stride
        self.h_ids = torch.empty((max_positions,), dtype=torch.int32).pin_memory()
        self.h_counts = torch.empty((self.m,), dtype=torch.int64).pin_memory()
        self.h_stage = [torch.empty((cap, stride), dtype=torch.uint8).pin_memory() for _ in range(2)]
        self.d_stage = [torch.empty((cap, stride), dtype=torch.uint8, device=dev) for _ in range(2)]
        self.side = torch.cuda.Stream(device=dev)
        self.cap = cap

    def lookup(self, input_ids: torch.Tensor, out: Optional[torch.Tensor] = None):
        ids = self.index._check_ids(input_ids)
        B, L = ids.shape
        dev, D = self.index.device, self.table.dim
        if out is None:
            out = torch.empty((B, L, D), dtype=self.base.dtype, device=dev)
        fid_all, len_all = self.index.lookup(ids)
        main = torch.cuda.current_stream(dev)
        rows_per = (B + self.m - 1) // self.m
        bounds = [(b0, min(B, b0 + rows_per)) for b0 in range(0, B, rows_per)]
        # matched row numbers of every micro-batch, compacted, to the host in one go (the tier's one host sync)
        flat = fid_all.reshape(-1)
        hit_all = flat >= 0
        rows_all = flat[hit_all]
        counts = torch.stack([hit_all[b0 * L:b1 * L].sum() for b0, b1 in bounds])
        if rows_all.numel() > self.h_ids.numel():
            raise ValueError("batch larger than the staging buffers (max_positions)")
        self.h_ids[:rows_all.numel()].copy_(rows_all, non_blocking=True)
        self.h_counts[:len(bounds)].copy_(counts, non_blocking=True)
        main.synchronize()
        ks = self.h_counts[:len(bounds)].tolist()
        L_ = _lib.load()
        pending, off = [], 0
        for m, (b0, b1) in enumerate(bounds):
            k, buf = int(ks[m]), m % 2
            if k > self.cap:
                raise ValueError("micro-batch larger than the staging buffers")
            if m >= 2:
                pending[m - 2].synchronize()           # staging buffers of micro-batch m-2 are free again
            _lib.check(L_.scone_host_gather_rows(self.table.storage.data_ptr(), self.table.row_stride, self.table.num_rows,
                                                 self.h_ids.data_ptr() + 4 * off, k, self.h_stage[buf].data_ptr(), self.threads))
            off += k
            with torch.cuda.stream(self.side):
                self.d_stage[buf][:k].copy_(self.h_stage[buf][:k], non_blocking=True)
                copied = torch.cuda.Event()
                copied.record(self.side)
            main.wait_event(copied)
            hit = hit_all[b0 * L:b1 * L]
            slot = (torch.cumsum(hit, 0, dtype=torch.int32) - 1).masked_fill_(~hit, -1)
            view = self.table.view_of(self.d_stage[buf][:max(k, 1)])
            embed_gather(view, self.base, ids[b0:b1], slot.view(b1 - b0, L), out=out[b0:b1])
            done = torch.cuda.Event()
            done.record(main)
            pending.append(done)
        return out, fid_all, len_all
