"""Host-fed lookups, multi-buffered: ids arrive in pinned HOST memory, match results go back to the host, and the
copies of neighbouring batches overlap the kernel.  Thin wrapper over the native ``scone_pipeline_*`` runtime
(``csrc/pipeline.cu``: a copy-in stream, two alternating compute streams, a copy-out stream, per-slot events, ``cudaMemcpyAsync`` in and out around the fused kernel).

    pipe = HostPipeline(index, table, base_emb, batch_shape=(B, L))
    for h_ids in batches:                       # pinned int64 [B, L] tensors
        done = pipe.submit(h_ids)               # the oldest finished batch, or None while the pipeline fills
    rest = pipe.flush()

Each result is ``(embeds [B, L, D] on the device, fgram_id int32 [B, L] pinned host, match_len uint8 [B, L] pinned host)``
and stays valid until the next ``submit`` (four slots by default: up to three batches in flight, one held by the caller;
measured on config 2: 2 slots 0.70, 3 slots 1.31, 4 slots 1.36 G tokens/s).  A result handed back by ``submit`` /
``flush`` is complete -- its copy-out, which follows the kernel, has been waited for -- so the embeddings can be consumed
on any stream without further synchronisation.
"""

from __future__ import annotations

import ctypes as C
from typing import Optional, Tuple

import torch

from . import _lib
from .index import FGramIndex
from .table import _OUT, CacheTable


class HostPipeline:
    def __init__(self, index: FGramIndex, table: CacheTable, base_emb: torch.Tensor, batch_shape: Tuple[int, int],
                 pos_emb: Optional[torch.Tensor] = None, slots: int = 4):
        B, L = batch_shape
        dev = index.device
        if base_emb.dtype not in (torch.bfloat16, torch.float16) or base_emb.device != dev or not base_emb.is_contiguous():
            raise ValueError("base_emb must be a contiguous bf16/fp16 tensor on the index device")
        self.index, self.table, self.base, self.pos = index, table, base_emb, pos_emb
        self.B, self.L, self.n = B, L, slots
        T = B * L
        self.d_ids = [torch.empty((B, L), dtype=torch.int64, device=dev) for _ in range(slots)]
        self.out = [torch.empty((B, L, table.dim), dtype=base_emb.dtype, device=dev) for _ in range(slots)]
        self.d_meta = [torch.empty((5 * T,), dtype=torch.uint8, device=dev) for _ in range(slots)]
        self.h_meta = [torch.empty((5 * T,), dtype=torch.uint8).pin_memory() for _ in range(slots)]
        self.h_id = [m[:4 * T].view(torch.int32).view(B, L) for m in self.h_meta]
        self.h_len = [m[4 * T:].view(B, L) for m in self.h_meta]
        self.status = torch.zeros((1,), dtype=torch.int32, device=dev)
        arr = lambda ts: (C.c_void_p * slots)(*[t.data_ptr() for t in ts])
        self._h = C.c_void_p()
        torch.cuda.synchronize(dev)
        with torch.cuda.device(dev):
            _lib.check(_lib.load().scone_pipeline_create(
                index.handle, C.byref(table.desc), base_emb.data_ptr(), base_emb.shape[0],
                pos_emb.data_ptr() if pos_emb is not None else None, B, L, _OUT[base_emb.dtype], slots,
                arr(self.d_ids), arr(self.out), arr(self.d_meta), arr(self.h_meta), self.status.data_ptr(), C.byref(self._h)))
        self.inflight = []
        self._slot = C.c_int32()

    def _result(self, slot: int):
        _lib.check(_lib.load().scone_pipeline_wait(self._h, slot))
        return self.out[slot], self.h_id[slot], self.h_len[slot]

    def submit(self, h_ids: torch.Tensor):
        """Enqueue one batch (pinned int64 [B, L]); returns the oldest finished result once ``slots - 1`` batches are in flight.
        The ids are read by an asynchronous copy: ``h_ids`` must stay untouched until this batch's result has been handed back."""
        if not h_ids.is_pinned() or h_ids.dtype != torch.int64 or h_ids.numel() != self.B * self.L or not h_ids.is_contiguous():
            raise ValueError("h_ids must be a contiguous pinned int64 host tensor of the pipeline's batch shape")
        _lib.check(_lib.load().scone_pipeline_submit(self._h, h_ids.data_ptr(), C.byref(self._slot)))
        self.inflight.append(int(self._slot.value))
        if len(self.inflight) >= max(1, self.n - 1):
            return self._result(self.inflight.pop(0))
        return None

    def follow(self, stream: Optional[torch.cuda.Stream] = None) -> None:
        """Order every batch submitted from now on behind the work already enqueued on ``stream`` (default: the current
        stream).  The pipeline runs on its own streams and is ordered against the caller only at construction, so
        call this after updating anything it reads -- ``table.store`` / ``cache_embeddings`` / ``set_base_embedding``."""
        st = stream if stream is not None else torch.cuda.current_stream(self.index.device)
        _lib.check(_lib.load().scone_pipeline_follow(self._h, st.cuda_stream))

    def flush(self):
        """Results of everything still in flight, oldest first."""
        res = [self._result(s) for s in self.inflight]
        self.inflight = []
        return res

    def close(self) -> None:
        if getattr(self, "_h", None):
            _lib.load().scone_pipeline_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
