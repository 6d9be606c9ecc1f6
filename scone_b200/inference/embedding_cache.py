"""Drop-in ``EmbeddingCache`` (mirror of reference ``scone/inference/embedding_cache.py``).

Same constructor / methods as the reference class.  Rows live in a packed device table
(``scone_b200.table.CacheTable``) instead of a dict of fp32 numpy rows or an fp32 ``np.memmap``:
unquantised fp32 by default -- what the reference stores and returns, bit for bit (:84-91, :132-135) --
or FP16 / INT8 / INT4 with ``quant=``; ``use_memory_map=True`` (the reference's "table does not live in
RAM" switch, :69-103) selects the offloaded tier: pinned host memory read zero-copy by the GPU.

Added for the fused path: :meth:`set_base_embedding` and :meth:`lookup`, which run
match + gather + dequant + fallback for a whole ``[B, L]`` batch in one kernel.
"""

from __future__ import annotations

import os
from collections.abc import Mapping
from typing import Dict, List, Optional, Union

import numpy as np
import torch

from ..table import CacheTable, embed_forward
from ..tokenization.n_gram_extractor import NGramExtractor, _default_device

_FORMAT = "scone_b200.cache.v1"


class _RowsView(Mapping):
    """``cache.embeddings``: id -> fp32 numpy row, like the reference's Dict[int, np.ndarray] (:49)."""

    def __init__(self, cache: "EmbeddingCache"):
        self._c = cache

    def _ids(self) -> np.ndarray:
        c = self._c
        if c._present is None:
            return np.zeros((0,), dtype=np.int64)
        return torch.nonzero(c._present).flatten().cpu().numpy()

    def __len__(self) -> int:
        return 0 if self._c._present is None else int(self._c._present.sum().item())

    def __iter__(self):
        return iter(int(i) for i in self._ids())

    def __contains__(self, key) -> bool:
        c = self._c
        return c._present is not None and isinstance(key, (int, np.integer)) and 0 <= int(key) < c._present.numel() \
            and bool(c._present[int(key)].item())

    def __getitem__(self, key) -> np.ndarray:
        if key not in self:
            raise KeyError(key)
        return self._c._gather_rows(torch.tensor([int(key)], device=self._c.device)).cpu().numpy()[0]


class EmbeddingCache:
    """Cache for f-gram embeddings (reference attributes: n_gram_extractor, embedding_dim, cache_dir,
    use_memory_map, embeddings, memory_mapped_embeddings)."""

    def __init__(
        self,
        n_gram_extractor: NGramExtractor,
        embedding_dim: int,
        cache_dir: Optional[str] = None,
        use_memory_map: bool = False,
        *,
        quant: str = "fp32",
        group_size: int = 128,
        out_dtype: torch.dtype = torch.bfloat16,
        device: Optional[torch.device] = None,
        tier: Optional[str] = None,
    ) -> None:
        """``quant``: "fp32" (default: rows kept exactly as given, like the reference) | "fp16" | "int8" | "int4"
        (``group_size`` for int4); ``tier``: "hbm" (default), "host" (pinned host
        memory, also selected by ``use_memory_map=True``) or "sharded" (rows split by id % world over the default
        process group and read over NVLink; needs ``torch.distributed`` initialised with NCCL)."""
        self.n_gram_extractor = n_gram_extractor
        self.embedding_dim = embedding_dim
        self.cache_dir = cache_dir
        self.use_memory_map = use_memory_map
        self.quant = quant
        self.group_size = group_size
        self.out_dtype = out_dtype
        self.tier = tier or ("host" if use_memory_map else "hbm")
        self._device = torch.device(device) if device is not None else None
        self._table: Optional[CacheTable] = None
        self._present: Optional[torch.Tensor] = None
        self._base_emb: Optional[torch.Tensor] = None
        self._pos_emb: Optional[torch.Tensor] = None
        self._status: Optional[torch.Tensor] = None
        self._sharded = None
        self._missing: Optional[int] = None     # cached: number of vocabulary rows never stored (None = not counted yet)
        if quant not in ("fp32", "fp16", "int8", "int4"):
            raise ValueError("quant must be 'fp32', 'fp16', 'int8' or 'int4'")
        if self.tier not in ("hbm", "host", "sharded"):
            raise ValueError("tier must be 'hbm', 'host' or 'sharded'")
        if self.cache_dir is not None and not os.path.exists(self.cache_dir):
            os.makedirs(self.cache_dir)

    # ---- plumbing -----------------------------------------------------------------------------------
    @property
    def device(self) -> torch.device:
        if self._device is None:
            self._device = _default_device()
        if self._device.type != "cuda":
            raise ValueError("scone_b200 has no CPU path: device must be CUDA")
        if self._device.index is None:
            self._device = torch.device("cuda", torch.cuda.current_device())
        return self._device

    @property
    def table(self) -> CacheTable:
        if self._table is None:
            if self.use_memory_map and self.cache_dir is None:
                raise ValueError("Cache directory must be provided for memory mapping")     # reference :72-73
            n = len(self.n_gram_extractor)
            if self.tier == "sharded":
                # rows split by id % world over the default process group, shards mapped by every rank (NVLink)
                from ..sharded import PeerShardedTable
                self._sharded = PeerShardedTable(n, self.embedding_dim, self.quant, self.group_size, self.device)
                self._table = self._sharded.local
            else:
                self._table = CacheTable(n, self.embedding_dim, self.quant, self.group_size, self.device, self.tier)
            self._present = torch.zeros((n,), dtype=torch.bool, device=self.device)
        return self._table

    @property
    def embeddings(self) -> Mapping:
        return _RowsView(self)

    @property
    def memory_mapped_embeddings(self) -> Optional[np.ndarray]:
        """Host-resident payload [N, D] (fp16 / int8) or [N, D/2] (int4 nibbles) when the table is offloaded."""
        if self._table is None or self.tier != "host":
            return None
        st = self._table.storage.numpy()
        D = self.embedding_dim
        if self.quant == "fp32":
            return st[:, :4 * D].view(np.float32)      # the reference's own memmap: [N, D] float32 (:84-91)
        if self.quant == "fp16":
            return st[:, :2 * D].view(np.float16)
        if self.quant == "int8":
            return st[:, :D].view(np.int8)
        return st[:, :D // 2]

    # ---- reference API --------------------------------------------------------------------------------
    def cache_embeddings(self, f_gram_ids: Union[List[int], Dict[int, torch.Tensor]],
                         embeddings: Optional[torch.Tensor] = None, verbose: bool = True,
                         projection: Optional[torch.Tensor] = None) -> None:
        """Store rows (reference :56-111).  Rows are quantised on the GPU, in chunks, straight into the table.

        ``projection`` ([embedding_dim, H_f], the ``nn.Linear.weight`` of the reference's bias-free ``f_gram_projection``,
        ``language_model.py:172-176``): ``embeddings`` is then ``[k, H_f]`` -- the f-gram model's own output -- and the table
        receives ``embeddings @ projection.T``, computed on the tensor cores with the quantise-and-store as the GEMM's
        epilogue, so that ``lookup`` serves rows already in the model's hidden size (the reference projects on every
        forward pass, ``:236``).

        Accepts the reference signature ``(f_gram_ids, embeddings[k, D])`` and also the ``{id: row}`` dict
        that the reference's own tests and scripts pass (tests/test_embedding_cache.py:78).
        """
        if isinstance(f_gram_ids, dict):
            items = list(f_gram_ids.items())
            ids = [int(k) for k, _ in items]
            embeddings = torch.stack([torch.as_tensor(v, dtype=torch.float32).cpu() for _, v in items]) if items \
                else torch.zeros((0, self.embedding_dim))
        elif isinstance(f_gram_ids, (torch.Tensor, np.ndarray, range)):
            ids = f_gram_ids                       # bulk form: no per-id Python objects (10^7-row tables)
        else:
            ids = [int(i) for i in f_gram_ids]
        if embeddings is None:
            raise ValueError("embeddings required")
        embeddings = torch.as_tensor(embeddings)
        width = self.embedding_dim if projection is None else projection.shape[1]
        if projection is not None and (projection.dim() != 2 or projection.shape[0] != self.embedding_dim):
            raise ValueError(f"projection must be [{self.embedding_dim}, H_f]")
        if embeddings.dim() != 2 or embeddings.shape[1] != width or embeddings.shape[0] != len(ids):
            raise ValueError(f"embeddings must be [{len(ids)}, {width}]")
        table = self.table
        if isinstance(ids, range):
            id_t = torch.arange(ids.start, ids.stop, ids.step, dtype=torch.int64, device=self.device)
        else:
            id_t = torch.as_tensor(ids, dtype=torch.int64).to(self.device)
        if len(ids) and (int(id_t.min()) < 0 or int(id_t.max()) >= len(self.n_gram_extractor)):
            raise IndexError("f-gram id out of range")                                     # memmap backend: IndexError
        self._missing = None
        if projection is not None:
            if self.tier != "hbm":
                raise ValueError("projection folding is available for the hbm tier")
            chunk = max(1, (256 << 20) // (4 * width))
            proj = projection.detach().to(device=self.device, dtype=torch.bfloat16)
            for s in range(0, len(ids), chunk):
                table.store_projected(embeddings[s:s + chunk].to(self.device, non_blocking=True), proj, id_t[s:s + chunk])
            self._present[id_t] = True
            return
        if self.tier == "sharded":
            # every rank may be handed every row; each keeps the ones it owns.  Call publish() after the last store.
            self._sharded.store_owned(embeddings.to(torch.float32), id_t)
            self._present[id_t] = True
            return
        chunk = max(1, (256 << 20) // (4 * self.embedding_dim))
        for s in range(0, len(ids), chunk):
            rows = embeddings[s:s + chunk].to(device=self.device, dtype=torch.float32, non_blocking=True)
            table.store(rows, id_t[s:s + chunk])
        self._present[id_t] = True

    def get_embeddings(self, f_gram_ids: List[int], device: Optional[torch.device] = None) -> torch.Tensor:
        """``table[ids]`` as fp32 [k, D] (reference :113-147): the stored rows, dequantised on the GPU.

        Like the reference the result is a fresh fp32 tensor, left on the CPU unless ``device`` is given.
        """
        if self._table is None:
            if self.use_memory_map:
                raise ValueError("Memory-mapped embeddings not initialized")               # reference :129-130
            if len(f_gram_ids):
                raise KeyError(f_gram_ids[0])
            return torch.zeros((0, self.embedding_dim))
        ids = torch.as_tensor(list(f_gram_ids), dtype=torch.int64, device=self.device)
        if ids.numel():
            bad = (ids < 0) | (ids >= self._present.numel())
            if bool(bad.any()):
                raise (IndexError if self.use_memory_map else KeyError)(int(ids[bad][0]))
            if not self.use_memory_map and not bool(self._present[ids].all()):
                raise KeyError(int(ids[~self._present[ids]][0]))                           # dict backend: KeyError (:139)
        out = self._gather_rows(ids)
        if device is None:
            return out.cpu()
        return out.to(device)

    def _gather_rows(self, ids: torch.Tensor) -> torch.Tensor:
        """dequant(table[ids]) as fp32 [k, D] for GLOBAL f-gram ids.  On the sharded tier row ``id`` lives on rank
        ``id % W`` at local row ``id // W``: each owner's rows are read through its peer-mapped shard (NVLink), so the
        reference-API methods see the whole table from every rank (the peers must have called :meth:`publish`)."""
        if self.tier != "sharded":
            return self._table.gather(ids, torch.float32)
        sh = self._sharded
        out = torch.empty((ids.numel(), self.embedding_dim), dtype=torch.float32, device=self.device)
        owner = ids % sh.world
        for r in range(sh.world):
            sel = torch.nonzero(owner == r).flatten()
            if sel.numel():
                out[sel] = sh.shard_view(r).gather(torch.div(ids[sel], sh.world, rounding_mode="floor"), torch.float32)
        return out

    def get_token_embeddings(self, token_ids: List[int], device: Optional[torch.device] = None) -> Dict[int, torch.Tensor]:
        """pos -> [k_pos, D] rows of all f-grams containing the position (reference :149-181); one match_all
        launch and one gather launch for the whole sequence."""
        token_ids = list(token_ids)
        L = len(token_ids)
        if L == 0 or len(self.n_gram_extractor) == 0:
            return {}
        index = self.n_gram_extractor.device_index(self.device)
        ids = torch.tensor([token_ids], dtype=torch.long, device=self.device)
        all_ids = index.match_all(ids)[0].cpu().numpy()                                    # [L, max_n]
        per_pos: Dict[int, List[int]] = {i: [] for i in range(L)}
        for n in range(1, min(index.max_n + 1, L + 1)):
            for e in np.flatnonzero(all_ids[:, n - 1] >= 0):
                for j in range(int(e) - n + 1, int(e) + 1):
                    per_pos[j].append(int(all_ids[e, n - 1]))
        flat = [g for i in range(L) for g in per_pos[i]]
        if not flat:
            return {}
        rows = self.get_embeddings(flat, device if device is not None else None)
        out, o = {}, 0
        for i in range(L):
            k = len(per_pos[i])
            if k:
                out[i] = rows[o:o + k]
                o += k
        return out

    # ---- the fused path -----------------------------------------------------------------------------------
    def set_base_embedding(self, weight: torch.Tensor, position_weight: Optional[torch.Tensor] = None) -> None:
        """Fallback rows = the model's token embedding ``wte.weight`` [V, D] (reference language_model.py:239),
        kept on the device in the output dtype; optional ``wpe.weight`` fuses the position add (:253-254)."""
        if weight.dim() != 2 or weight.shape[1] != self.embedding_dim:
            raise ValueError(f"base embedding must be [V, {self.embedding_dim}]")
        self._base_emb = weight.detach().to(device=self.device, dtype=self.out_dtype).contiguous()
        self._pos_emb = None if position_weight is None else \
            position_weight.detach().to(device=self.device, dtype=self.out_dtype).contiguous()

    def _require_all_rows(self) -> None:
        """The reference raises KeyError when a matched f-gram has no cached row (``embedding_cache.py:139``).  The fused
        kernels do not consult the presence mask per position, so the check is made once per change of the table: every
        vocabulary row must have been stored before the first ``lookup``."""
        if self._missing is None:
            self.table
            self._missing = int((~self._present).sum().item())
        if self._missing:
            raise KeyError(int(torch.nonzero(~self._present).flatten()[0].item()))

    def lookup(self, input_ids: torch.Tensor, out: Optional[torch.Tensor] = None, add_positions: bool = False,
               combine: str = "replace", inputs_stable: bool = False, strict: bool = True,
               out_id: Optional[torch.Tensor] = None, out_len: Optional[torch.Tensor] = None):
        """``input_ids`` long [B, L] on the GPU -> (embeds [B, L, D] out_dtype, fgram_id int32 [B, L], match_len uint8 [B, L]).

        embeds[b, i] = dequant(row of the longest f-gram ending at i) or base_emb[input_ids[b, i]].
        ``combine="add"``: the reference code's ``wte(input_ids) + f_gram_embeddings`` (``language_model.py:239-243``)
        instead -- the row is added to the token embedding (hbm / host tiers).  Asynchronous on the current stream.
        ``strict`` (default): raise ``KeyError`` like the reference if some vocabulary f-gram has no cached row yet
        (checked once after each ``cache_embeddings``; ``strict=False`` serves such f-grams from zero-filled storage).
        ``inputs_stable``: see :func:`scone_b200.table.embed_forward`.
        """
        if self._base_emb is None:
            raise RuntimeError("call set_base_embedding(wte.weight) first: misses fall back to the token embedding")
        if add_positions and self._pos_emb is None:
            raise RuntimeError("add_positions=True needs set_base_embedding(..., position_weight=wpe.weight)")
        index = self.n_gram_extractor.device_index(self.device)
        if strict:
            self._require_all_rows()
        if self._status is None:
            self._status = torch.zeros((1,), dtype=torch.int32, device=self.device)
        if self.tier == "sharded":
            if combine != "replace":
                raise ValueError("combine='add' is not available on the sharded tier")
            from ..sharded import embed_forward_sharded
            self.table
            return embed_forward_sharded(index, self._sharded, self._base_emb, input_ids,
                                         self._pos_emb if add_positions else None, out, self._status, out_id, out_len)
        return embed_forward(index, self.table, self._base_emb, input_ids, self._pos_emb if add_positions else None, out,
                             self._status, out_id=out_id, out_len=out_len, combine=combine, inputs_stable=inputs_stable)

    def host_pipeline(self, batch_shape, add_positions: bool = False, slots: int = 4, strict: bool = True):
        """A :class:`scone_b200.HostPipeline` over this cache for callers whose ids arrive in pinned HOST memory
        (the reference tokenises on the host, ``engine.py:222-233``): copies in and out overlap the kernel."""
        from ..pipeline import HostPipeline
        if self._base_emb is None:
            raise RuntimeError("call set_base_embedding(wte.weight) first")
        if self.tier == "sharded":
            raise ValueError("host_pipeline is for the hbm / host tiers")
        if strict:
            self._require_all_rows()
        index = self.n_gram_extractor.device_index(self.device)
        return HostPipeline(index, self.table, self._base_emb, tuple(batch_shape),
                            pos_emb=self._pos_emb if add_positions else None, slots=slots)

    def publish(self) -> None:
        """Sharded tier: make this rank's rows visible to its peers (collective; call once after the last store)."""
        if self.tier == "sharded":
            self.table
            self._sharded.publish()

    def _not_on_sharded(self, what: str) -> None:
        # a rank holds only ceil(N / W) rows (local row = id // W): persisting or averaging "the table" from one rank
        # would silently use the wrong rows
        if self.tier == "sharded":
            raise NotImplementedError(f"EmbeddingCache.{what} is not available on the row-sharded tier: every rank holds "
                                      "only its own rows (save each shard from an hbm-tier cache before sharding)")

    def status(self) -> int:
        """Sticky device status bits (synchronises): bit 0 = some missed token id was outside the base table."""
        return 0 if self._status is None else int(self._status.item())

    # ---- fixed binary format -------------------------------------------------------------
    _MAGIC = b"SCONEB2\x00"

    def save_binary(self, path: str) -> None:
        """Header-carrying flat file: 64-byte header, then num_rows x row_stride row bytes, then num_rows presence bytes.
        Unlike the reference's raw ``np.memmap`` (``embedding_cache.py:84-89``) it can be reloaded (and memory-mapped)."""
        import struct
        t = self.table
        self._not_on_sharded("save_binary")
        hdr = struct.pack("<8sIIIIQIQ", self._MAGIC, 1, {"fp16": 0, "int8": 1, "int4": 2, "fp32": 3}[self.quant], self.embedding_dim,
                          self.group_size, t.row_stride, t.scale_offset, t.num_rows)
        with open(path, "wb") as f:
            f.write(hdr.ljust(64, b"\x00"))
            st = t.storage
            step = max(1, (256 << 20) // max(1, t.row_stride))
            for s0 in range(0, t.num_rows, step):
                f.write(st[s0:s0 + step].cpu().numpy().tobytes())
            f.write(self._present.cpu().numpy().astype(np.uint8).tobytes())

    @classmethod
    def load_binary(cls, path: str, n_gram_extractor: NGramExtractor, cache_dir: Optional[str] = None,
                    use_memory_map: bool = False, **kwargs) -> "EmbeddingCache":
        import struct
        with open(path, "rb") as f:
            raw = f.read(64)
        magic, ver, q, dim, group, stride, soff, nrows = struct.unpack("<8sIIIIQIQ", raw[:struct.calcsize("<8sIIIIQIQ")])
        if magic != cls._MAGIC or ver != 1:
            raise ValueError(f"{path}: not a scone_b200 cache file")
        cache = cls(n_gram_extractor, dim, cache_dir=cache_dir, use_memory_map=use_memory_map,
                    quant=["fp16", "int8", "int4", "fp32"][q], group_size=group, **kwargs)
        t = cache.table
        if (t.row_stride, t.scale_offset, t.num_rows) != (stride, soff, nrows):
            raise ValueError("file geometry does not match this vocabulary / build")
        mm = np.memmap(path, dtype=np.uint8, mode="r", offset=64, shape=(nrows, stride))
        step = max(1, (256 << 20) // max(1, stride))
        for s0 in range(0, nrows, step):
            t.storage[s0:s0 + step].copy_(torch.from_numpy(np.array(mm[s0:s0 + step])))
        present = np.fromfile(path, dtype=np.uint8, offset=64 + nrows * stride, count=nrows)
        cache._present.copy_(torch.from_numpy(present.astype(bool)))
        return cache

    @classmethod
    def from_reference_memmap(cls, mmap_path: str, n_gram_extractor: NGramExtractor, embedding_dim: int,
                              **kwargs) -> "EmbeddingCache":
        """Import the raw fp32 ``[N, D]`` file the reference's memmap backend writes to ``cache_dir/embeddings.npy``
        (``embedding_cache.py:76-91``; it has no ``.npy`` header, which is why the reference cannot reload it)."""
        n = len(n_gram_extractor)
        rows = np.memmap(mmap_path, dtype=np.float32, mode="r", shape=(n, embedding_dim))
        cache = cls(n_gram_extractor, embedding_dim, **kwargs)
        step = max(1, (256 << 20) // (4 * embedding_dim))
        for s0 in range(0, n, step):
            blk = torch.from_numpy(np.ascontiguousarray(rows[s0:s0 + step]))
            cache.cache_embeddings(list(range(s0, s0 + blk.shape[0])), blk, verbose=False)
        return cache

    def assemble_mean(self, input_ids: torch.Tensor, dtype: torch.dtype = torch.float32) -> torch.Tensor:
        """The reference engine's own tensor (``engine.py:235-259``): mean of the rows of all f-grams containing each
        position, zeros where none.  Optional mode; :meth:`lookup` is the Algorithm-2 path."""
        from ..table import embed_mean_forward
        self._not_on_sharded("assemble_mean")
        return embed_mean_forward(self.n_gram_extractor.device_index(self.device), self.table, input_ids, dtype)

    # ---- persistence (reference :183-243) ----------------------------------------------------------------------
    def save(self, path: str) -> None:
        """One ``.npy`` pickle like the reference, but carrying the packed table and its geometry
        (the reference's memmap backend cannot be reloaded: raw memmap written :84-89, ``np.load`` at :232)."""
        self._not_on_sharded("save")
        t = self.table
        present = self._present.cpu().numpy()
        np.save(path, {
            "format": _FORMAT, "use_memory_map": self.use_memory_map, "cache_dir": self.cache_dir,
            "embedding_dim": self.embedding_dim, "quant": self.quant, "group_size": self.group_size,
            "row_stride": t.row_stride, "scale_offset": t.scale_offset, "num_rows": t.num_rows,
            "storage": t.storage.cpu().numpy(), "present": present,
        }, allow_pickle=True)

    @classmethod
    def load(cls, path: str, n_gram_extractor: NGramExtractor, cache_dir: Optional[str] = None, **kwargs) -> "EmbeddingCache":
        """Load our format, or a dict-backend cache written by the reference (``{"embeddings": {id: fp32 row}}``)."""
        path = str(path)
        if not os.path.exists(path) and os.path.exists(path + ".npy"):
            path = path + ".npy"
        data = np.load(path, allow_pickle=True).item()
        if data.get("format") == _FORMAT:
            cache = cls(n_gram_extractor, data["embedding_dim"], cache_dir=cache_dir or data["cache_dir"],
                        use_memory_map=data["use_memory_map"], quant=data["quant"], group_size=data["group_size"], **kwargs)
            t = cache.table
            if (t.row_stride, t.num_rows) != (data["row_stride"], data["num_rows"]):
                raise ValueError("saved table geometry does not match this vocabulary")
            t.storage.copy_(torch.from_numpy(data["storage"]))
            cache._present.copy_(torch.from_numpy(data["present"]))
            return cache
        if data.get("use_memory_map"):
            raise ValueError("reference memmap caches carry no rows in the .npy file; load the raw fp32 file and "
                             "call cache_embeddings")
        cache = cls(n_gram_extractor, data["embedding_dim"], cache_dir=cache_dir, use_memory_map=False, **kwargs)
        emb = data["embeddings"]
        if emb:
            ids = sorted(emb)
            cache.cache_embeddings(ids, torch.from_numpy(np.stack([np.asarray(emb[i], dtype=np.float32) for i in ids])),
                                   verbose=False)
        return cache
