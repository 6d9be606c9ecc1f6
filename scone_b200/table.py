"""Packed cache table in HBM (or device-mapped pinned host memory) + the fused embed call."""

from __future__ import annotations

import ctypes as C
from typing import Optional, Tuple

import torch

from . import _lib
from .index import FGramIndex, _require_cuda, _stream_ptr

_OUT = {torch.bfloat16: _lib.OUT_BF16, torch.float16: _lib.OUT_FP16, torch.float32: _lib.OUT_FP32}


def table_layout(quant: str, dim: int, group: int = 128, align: int = 32) -> Tuple[int, int]:
    """(row_stride bytes, scale_offset bytes) of a stored row."""
    if quant not in _lib.QUANT:
        raise ValueError(f"quant must be one of {sorted(_lib.QUANT)}")
    rs, so = C.c_int64(), C.c_int32()
    _lib.check(_lib.load().scone_table_layout(_lib.QUANT[quant], dim, group, align, C.byref(rs), C.byref(so)))
    return int(rs.value), int(so.value)


class CacheTable:
    """N rows of D elements stored as FP16 / INT8 (per-row scale) / INT4 (per-group fp16 scales).

    Row r is the embedding of f-gram id r (reference ``scone/inference/embedding_cache.py:77,99``).
    ``storage`` is a uint8 tensor [N, row_stride]: on the GPU (tier "hbm") or pinned host memory
    mapped into the device address space (tier "host").
    """

    def __init__(self, num_rows: int, dim: int, quant: str = "fp16", group: int = 128, device="cuda", tier: str = "hbm",
                 align: int = 32, storage: Optional[torch.Tensor] = None):
        self.num_rows, self.dim, self.quant, self.group, self.tier = int(num_rows), int(dim), quant, int(group), tier
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise ValueError("CacheTable lives on a CUDA device: scone_b200 has no CPU path")
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        self.row_stride, self.scale_offset = table_layout(quant, dim, group, align)
        if storage is not None:
            if storage.dtype != torch.uint8 or tuple(storage.shape) != (self.num_rows, self.row_stride):
                raise ValueError("storage must be uint8 [num_rows, row_stride]")
            self.storage = storage
        elif tier == "hbm":
            self.storage = torch.zeros((self.num_rows, self.row_stride), dtype=torch.uint8, device=self.device)
        elif tier == "host":
            self.storage = torch.zeros((self.num_rows, self.row_stride), dtype=torch.uint8, pin_memory=True)
        else:
            raise ValueError("tier must be 'hbm' or 'host'")
        if tier == "hbm":
            _require_cuda(self.storage, "storage")
            self._dev_ptr = self.storage.data_ptr()
        else:
            if not self.storage.is_pinned():
                raise ValueError("host-tier storage must be pinned")
            # pinned allocations are mapped under UVA: the host pointer is valid on the device
            self._dev_ptr = self.storage.data_ptr()
        self.desc = _lib.TableDesc(self._dev_ptr, self.row_stride, self.num_rows, _lib.QUANT[quant], self.dim, self.group,
                                   self.scale_offset)

    @property
    def bytes(self) -> int:
        return self.num_rows * self.row_stride

    def store(self, rows_fp32: torch.Tensor, row_ids: Optional[torch.Tensor] = None, row_base: int = 0) -> None:
        """Quantise fp32 rows [k, D] on the GPU and write them at ``row_ids`` (or row_base..row_base+k)."""
        _require_cuda(rows_fp32, "rows")
        rows = rows_fp32.to(torch.float32).contiguous()
        if rows.dim() != 2 or rows.shape[1] != self.dim:
            raise ValueError(f"rows must be [k, {self.dim}]")
        k = rows.shape[0]
        ids_ptr = None
        if row_ids is not None:
            row_ids = row_ids.to(device=self.device, dtype=torch.int64).contiguous()
            if row_ids.numel() != k:
                raise ValueError("row_ids and rows disagree on k")
            if k and (int(row_ids.min()) < 0 or int(row_ids.max()) >= self.num_rows):
                raise IndexError("row id out of range")
            ids_ptr = row_ids.data_ptr()
        with torch.cuda.device(self.device):
            _lib.check(_lib.load().scone_table_store(C.byref(self.desc), rows.data_ptr(), ids_ptr, int(row_base), k,
                                                     _stream_ptr(self.device)))

    def gather(self, row_ids: torch.Tensor, dtype: torch.dtype = torch.float32) -> torch.Tensor:
        """dequant(table[row_ids]) -> [k, D] (the reference's get_embeddings gather)."""
        row_ids = row_ids.to(device=self.device, dtype=torch.int64).contiguous()
        k = row_ids.numel()
        out = torch.empty((k, self.dim), dtype=dtype, device=self.device)
        with torch.cuda.device(self.device):
            _lib.check(_lib.load().scone_table_gather(C.byref(self.desc), row_ids.data_ptr(), k, out.data_ptr(), _OUT[dtype],
                                                      None, _stream_ptr(self.device)))
        return out


    def gather_packed(self, row_ids: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """Stored bytes of rows ``row_ids`` (int32) -> uint8 [k, row_stride]; no dequantisation (sharded tier, owner side)."""
        row_ids = row_ids.to(device=self.device, dtype=torch.int32).contiguous()
        k = row_ids.numel()
        if out is None:
            out = torch.empty((k, self.row_stride), dtype=torch.uint8, device=self.device)
        with torch.cuda.device(self.device):
            _lib.check(_lib.load().scone_table_gather_packed(C.byref(self.desc), row_ids.data_ptr(), k, out.data_ptr(), None,
                                                             _stream_ptr(self.device)))
        return out

    def view_of(self, storage: torch.Tensor) -> "CacheTable":
        """A table with this one's format over another [k, row_stride] uint8 buffer (e.g. rows received from a peer)."""
        return CacheTable(storage.shape[0], self.dim, self.quant, self.group, self.device, "hbm", storage=storage)


def embed_forward(index: FGramIndex, table: CacheTable, base_emb: torch.Tensor, input_ids: torch.Tensor,
                  pos_emb: Optional[torch.Tensor] = None, out: Optional[torch.Tensor] = None,
                  status: Optional[torch.Tensor] = None, want_ids: bool = True,
                  out_id: Optional[torch.Tensor] = None, out_len: Optional[torch.Tensor] = None, combine: str = "replace"):
    """The fused hot path.  Returns (embeds [B, L, D] in base_emb.dtype, fgram_id int32 [B, L], match_len uint8 [B, L]).

    out[b, i] = dequant(table[fgram_id[b, i]]) if an f-gram ends at (b, i) else base_emb[input_ids[b, i]]
    (+ pos_emb[i] when given).  ``combine="add"`` is the reference code's combine instead of Algorithm 2's replacement
    (``scone/models/language_model.py:239-243``): base_emb[input_ids[b, i]] + dequant(row) where an f-gram ends.
    Everything is enqueued on the current stream; nothing synchronises.
    """
    if combine not in ("replace", "add"):
        raise ValueError("combine must be 'replace' or 'add'")
    ids = index._check_ids(input_ids)
    B, L = ids.shape
    dev = index.device
    if base_emb.device != dev or table.device != dev:
        raise ValueError("index, table and base_emb must be on the same device")
    if base_emb.dtype not in (torch.bfloat16, torch.float16):
        raise ValueError("base_emb must be bf16 or fp16 (it defines the output dtype)")
    if base_emb.dim() != 2 or base_emb.shape[1] != table.dim or not base_emb.is_contiguous():
        raise ValueError(f"base_emb must be contiguous [V, {table.dim}]")
    if pos_emb is not None:
        if pos_emb.dtype != base_emb.dtype or pos_emb.device != dev or not pos_emb.is_contiguous() \
                or pos_emb.dim() != 2 or pos_emb.shape[1] != table.dim or pos_emb.shape[0] < L:
            raise ValueError(f"pos_emb must be contiguous [>= {L}, {table.dim}] {base_emb.dtype} on {dev}")
    if out is None:
        out = torch.empty((B, L, table.dim), dtype=base_emb.dtype, device=dev)
    elif out.dtype != base_emb.dtype or tuple(out.shape) != (B, L, table.dim) or not out.is_contiguous() or out.device != dev:
        raise ValueError("out must be contiguous [B, L, D] in base_emb.dtype")
    if want_ids:
        if out_id is None:
            out_id = torch.empty((B, L), dtype=torch.int32, device=dev)
        if out_len is None:
            out_len = torch.empty((B, L), dtype=torch.uint8, device=dev)
        if out_id.dtype != torch.int32 or out_len.dtype != torch.uint8 or out_id.numel() != B * L or out_len.numel() != B * L \
                or not out_id.is_contiguous() or not out_len.is_contiguous() or out_id.device != dev or out_len.device != dev:
            raise ValueError("out_id must be contiguous int32 [B, L] and out_len uint8 [B, L] on the index device")
    else:
        out_id = out_len = None
    with torch.cuda.device(dev):
        entry = _lib.load().scone_embed_forward_additive if combine == "add" else _lib.load().scone_embed_forward
        _lib.check(entry(
            index.handle, C.byref(table.desc), base_emb.data_ptr(), base_emb.shape[0],
            pos_emb.data_ptr() if pos_emb is not None else None, ids.data_ptr(), B, L, out.data_ptr(), _OUT[base_emb.dtype],
            out_id.data_ptr() if want_ids else None, out_len.data_ptr() if want_ids else None,
            status.data_ptr() if status is not None else None, _stream_ptr(dev)))
    return out, out_id, out_len


def embed_gather(table: CacheTable, base_emb: torch.Tensor, input_ids: torch.Tensor, fgram_id: torch.Tensor,
                 pos_emb: Optional[torch.Tensor] = None, out: Optional[torch.Tensor] = None,
                 status: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Second half of the path with ids already resolved (sharded / staged tiers)."""
    _require_cuda(input_ids, "input_ids")
    ids = input_ids.contiguous()
    fid = fgram_id.contiguous()
    if ids.dtype != torch.int64 or fid.dtype != torch.int32 or ids.shape != fid.shape:
        raise ValueError("input_ids (long) and fgram_id (int32) must have the same shape")
    dev = table.device
    T = ids.numel()
    L = ids.shape[-1] if ids.dim() >= 1 else 1
    if out is None:
        out = torch.empty(tuple(ids.shape) + (table.dim,), dtype=base_emb.dtype, device=dev)
    with torch.cuda.device(dev):
        _lib.check(_lib.load().scone_embed_gather(
            C.byref(table.desc), base_emb.data_ptr(), base_emb.shape[0], pos_emb.data_ptr() if pos_emb is not None else None,
            L, ids.data_ptr(), fid.data_ptr(), T, out.data_ptr(), _OUT[base_emb.dtype],
            status.data_ptr() if status is not None else None, _stream_ptr(dev)))
    return out


def embed_mean_forward(index: FGramIndex, table: CacheTable, input_ids: torch.Tensor, dtype: torch.dtype = torch.float32) -> torch.Tensor:
    """Reference-CODE semantics (``scone/inference/engine.py:235-259``): per position the mean of the rows of all f-grams
    containing it, zeros where none -- the tensor the reference passes as ``f_gram_embeddings``.  [B, L, D] in ``dtype``."""
    ids = index._check_ids(input_ids)
    B, L = ids.shape
    dev = index.device
    work = torch.empty((B, L, index.max_n), dtype=torch.int32, device=dev)
    out = torch.empty((B, L, table.dim), dtype=dtype, device=dev)
    with torch.cuda.device(dev):
        _lib.check(_lib.load().scone_embed_mean_forward(index.handle, C.byref(table.desc), ids.data_ptr(), B, L, work.data_ptr(),
                                                        out.data_ptr(), _OUT[dtype], _stream_ptr(dev)))
    return out
