"""Host-fed lookups, double-buffered: ids arrive in pinned HOST memory, match results go back to the host, and the
copies of neighbouring batches overlap the kernel.

    pipe = HostPipeline(cache_or_parts, batch_shape=(B, L))
    for h_ids in batches:                       # pinned int64 [B, L] tensors
        done = pipe.submit(h_ids)               # returns the PREVIOUS batch's result (or None for the first)
    last = pipe.flush()

Each result is ``(embeds [B, L, D] on the device, fgram_id int32 [B, L] pinned host, match_len uint8 [B, L] pinned host)``
and stays valid until the next ``submit`` (three slots: two batches in flight, one held by the caller).  Three streams:
copy-in, compute (the fused kernel), copy-out; events chain them per slot, so batch k+1's H2D and batch k-1's D2H run
under batch k's kernel.
"""

from __future__ import annotations

from typing import Optional, Tuple

import torch

from .index import FGramIndex
from .table import CacheTable, embed_forward


class HostPipeline:
    def __init__(self, index: FGramIndex, table: CacheTable, base_emb: torch.Tensor, batch_shape: Tuple[int, int],
                 pos_emb: Optional[torch.Tensor] = None, slots: int = 3):
        B, L = batch_shape
        dev = index.device
        self.index, self.table, self.base, self.pos = index, table, base_emb, pos_emb
        self.n = slots
        self.s_in, self.s_run, self.s_out = (torch.cuda.Stream(device=dev) for _ in range(3))
        mk = lambda *shape, dtype: torch.empty(shape, dtype=dtype, device=dev)
        self.d_ids = [mk(B, L, dtype=torch.int64) for _ in range(slots)]
        self.out = [mk(B, L, table.dim, dtype=base_emb.dtype) for _ in range(slots)]
        self.d_id = [mk(B, L, dtype=torch.int32) for _ in range(slots)]
        self.d_len = [mk(B, L, dtype=torch.uint8) for _ in range(slots)]
        self.h_id = [torch.empty((B, L), dtype=torch.int32).pin_memory() for _ in range(slots)]
        self.h_len = [torch.empty((B, L), dtype=torch.uint8).pin_memory() for _ in range(slots)]
        self.status = torch.zeros((1,), dtype=torch.int32, device=dev)
        self.ev_in = [torch.cuda.Event() for _ in range(slots)]
        self.ev_run = [torch.cuda.Event() for _ in range(slots)]
        self.ev_out = [torch.cuda.Event() for _ in range(slots)]
        self.k = 0
        self.inflight = []

    def _result(self, slot):
        self.ev_out[slot].synchronize()
        return self.out[slot], self.h_id[slot], self.h_len[slot]

    def submit(self, h_ids: torch.Tensor):
        """Enqueue one batch (pinned int64 [B, L]); returns the oldest finished result once the pipeline is full."""
        if not h_ids.is_pinned() or h_ids.dtype != torch.int64:
            raise ValueError("h_ids must be a pinned int64 host tensor")
        slot = self.k % self.n
        with torch.cuda.stream(self.s_in):
            self.s_in.wait_event(self.ev_run[slot])          # previous kernel on this slot has consumed d_ids
            self.d_ids[slot].copy_(h_ids, non_blocking=True)
            self.ev_in[slot].record(self.s_in)
        with torch.cuda.stream(self.s_run):
            self.s_run.wait_event(self.ev_in[slot])
            self.s_run.wait_event(self.ev_out[slot])         # previous results of this slot have left the device
            embed_forward(self.index, self.table, self.base, self.d_ids[slot], pos_emb=self.pos, out=self.out[slot],
                          status=self.status, out_id=self.d_id[slot], out_len=self.d_len[slot])
            self.ev_run[slot].record(self.s_run)
        with torch.cuda.stream(self.s_out):
            self.s_out.wait_event(self.ev_run[slot])
            self.h_id[slot].copy_(self.d_id[slot], non_blocking=True)
            self.h_len[slot].copy_(self.d_len[slot], non_blocking=True)
            self.ev_out[slot].record(self.s_out)
        self.inflight.append(slot)
        self.k += 1
        # up to n - 1 batches are in flight while we wait for the oldest; its slot is not reused before the submit after next
        if len(self.inflight) >= max(1, self.n - 1):
            return self._result(self.inflight.pop(0))
        return None

    def flush(self):
        """Results of everything still in flight, oldest first."""
        res = [self._result(s) for s in self.inflight]
        self.inflight = []
        return res
