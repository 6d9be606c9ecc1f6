"""Drop-in ``NGramExtractor`` (mirror of reference ``scone/tokenization/n_gram_extractor.py``).

Same constructor, attributes and methods as the reference class; the membership tests that
dominate ``get_token_f_grams`` (reference :119-124) run on the GPU through the f-gram index,
and a batched ``lookup`` is added for the fused path.  The vocabulary's source of truth is a
pair of flat arrays (tokens int32 [N, max_n] padded with -1, lens uint8 [N], id = row); the
reference's three Python maps are materialised lazily so 10^7..10^8-entry vocabularies do not
pay ~360 B/f-gram of Python objects unless someone asks for them.
"""

from __future__ import annotations

from typing import Dict, Iterable, List, Optional, Sequence, Set, Tuple

import numpy as np
import torch

from ..index import FGramIndex

FGram = Tuple[int, ...]


def _default_device() -> torch.device:
    if not torch.cuda.is_available():
        raise RuntimeError("scone_b200 needs a CUDA device (B200, sm_100a): there is no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device())


class NGramExtractor:
    """Extracts n-grams and identifies frequent n-grams (f-grams).

    Attributes (as in the reference, n_gram_extractor.py:17-24):
        max_n, min_freq, max_f_grams, f_grams, f_gram_to_id, id_to_f_gram
    """

    def __init__(self, max_n: int = 3, min_freq: int = 100, max_f_grams: int = 10_000_000,
                 device: Optional[torch.device] = None) -> None:
        self.max_n = max_n
        self.min_freq = min_freq
        self.max_f_grams = max_f_grams
        self.device = device
        self._tokens = np.zeros((0, max_n), dtype=np.int32)
        self._lens = np.zeros((0,), dtype=np.uint8)
        self._maps = None            # (set, dict, dict) built on demand
        self._index: Optional[FGramIndex] = None

    # ---- vocabulary as flat arrays ------------------------------------------------------------
    @classmethod
    def from_arrays(cls, vocab_tokens, vocab_lens, min_freq: int = 1, max_f_grams: Optional[int] = None,
                    device: Optional[torch.device] = None) -> "NGramExtractor":
        """Build from tokens int32 [N, max_n] (pad -1) and lens uint8 [N]; id = row number."""
        toks = np.ascontiguousarray(vocab_tokens, dtype=np.int32)
        lens = np.ascontiguousarray(vocab_lens, dtype=np.uint8)
        if toks.ndim != 2 or lens.shape != (toks.shape[0],):
            raise ValueError("vocab_tokens must be [N, max_n] and vocab_lens [N]")
        ex = cls(max_n=toks.shape[1], min_freq=min_freq, max_f_grams=max_f_grams or max(1, len(lens)), device=device)
        ex._tokens, ex._lens = toks, lens
        return ex

    def vocab_arrays(self) -> Tuple[np.ndarray, np.ndarray]:
        return self._tokens, self._lens

    def __len__(self) -> int:
        return int(self._lens.shape[0])

    def _set_vocab(self, grams_in_id_order: Sequence[FGram]) -> None:
        n = len(grams_in_id_order)
        width = max([self.max_n] + [len(g) for g in grams_in_id_order]) if n else self.max_n
        toks = np.full((n, width), -1, dtype=np.int32)
        lens = np.zeros((n,), dtype=np.uint8)
        for i, g in enumerate(grams_in_id_order):
            toks[i, :len(g)] = g
            lens[i] = len(g)
        self._tokens, self._lens = toks, lens
        self._maps = None
        self._drop_index()

    def _drop_index(self) -> None:
        if self._index is not None:
            self._index.close()
            self._index = None

    def _materialise(self):
        if self._maps is None:
            grams = [tuple(int(t) for t in self._tokens[i, :self._lens[i]]) for i in range(len(self._lens))]
            self._maps = (set(grams), {g: i for i, g in enumerate(grams)}, dict(enumerate(grams)))
        return self._maps

    @property
    def f_grams(self) -> Set[FGram]:
        return self._materialise()[0]

    @f_grams.setter
    def f_grams(self, value) -> None:
        # the reference lets callers assign the three maps (load() does, :159-165); ids come from f_gram_to_id
        if self._maps is None or set(value) != self._maps[0]:
            self._set_vocab(list(value))

    @property
    def f_gram_to_id(self) -> Dict[FGram, int]:
        return self._materialise()[1]

    @f_gram_to_id.setter
    def f_gram_to_id(self, mapping: Dict[FGram, int]) -> None:
        items = sorted(mapping.items(), key=lambda kv: kv[1])
        if [v for _, v in items] != list(range(len(items))):
            raise ValueError("f-gram ids must be exactly 0..N-1")
        self._set_vocab([tuple(int(t) for t in k) for k, _ in items])

    @property
    def id_to_f_gram(self) -> Dict[int, FGram]:
        return self._materialise()[2]

    @id_to_f_gram.setter
    def id_to_f_gram(self, mapping: Dict[int, FGram]) -> None:
        self.f_gram_to_id = {tuple(g): i for i, g in mapping.items()}

    # ---- n-gram enumeration (reference :46-70) ------------------------------------------------
    def extract_n_grams(self, token_ids: List[int], n: int) -> List[FGram]:
        return [tuple(token_ids[i:i + n]) for i in range(len(token_ids) - n + 1)]

    def extract_all_n_grams(self, token_ids: List[int]) -> List[FGram]:
        out: List[FGram] = []
        for n in range(1, min(self.max_n + 1, len(token_ids) + 1)):
            out.extend(self.extract_n_grams(token_ids, n))
        return out

    # ---- fit (reference :72-104) ----------------------------------------------------------------
    def fit(self, tokenized_texts: Iterable[Sequence[int]], verbose: bool = True) -> "NGramExtractor":
        """Identify the frequent n-grams of a corpus.

        Same result as the reference: count all n-grams (n = 1..max_n, never across texts), keep the
        ``max_f_grams`` most frequent (ties: first seen, in the reference's enumeration order text ->
        n -> start), THEN drop those below ``min_freq``; id = rank.  Counting is vectorised with numpy
        (sort + run-length) instead of a Python ``Counter``.
        """
        texts = [np.asarray(t, dtype=np.int64).ravel() for t in tokenized_texts]
        if any(t.size and (t.min() < 0 or t.max() > 0x7FFFFFFF) for t in texts):
            raise ValueError("token ids must be in [0, 2^31)")       # -1 is the padding value of the flat arrays
        max_n = self.max_n
        stride = max([len(t) for t in texts] + [1])
        uniq_rows, uniq_cnt, uniq_first = [], [], []
        for n in range(1, max_n + 1):
            wins, firsts = [], []
            for ti, t in enumerate(texts):
                if len(t) < n:
                    continue
                w = np.lib.stride_tricks.sliding_window_view(t, n)
                wins.append(w)
                # enumeration order of the reference: text, then n, then start
                firsts.append((ti * max_n + (n - 1)) * stride + np.arange(len(w), dtype=np.int64))
            if not wins:
                continue
            w = np.concatenate(wins)
            f = np.concatenate(firsts)
            order = np.lexsort([f] + [w[:, k] for k in range(n - 1, -1, -1)])
            ws, fs = w[order], f[order]
            new = np.ones(len(ws), dtype=bool)
            new[1:] = np.any(ws[1:] != ws[:-1], axis=1)
            starts = np.flatnonzero(new)
            cnt = np.diff(np.append(starts, len(ws)))
            rows = np.full((len(starts), max_n), -1, dtype=np.int64)
            rows[:, :n] = ws[starts]
            uniq_rows.append(rows)
            uniq_cnt.append(cnt)
            uniq_first.append(fs[starts])          # lexsort put the earliest occurrence first
        if uniq_rows:
            rows = np.concatenate(uniq_rows)
            cnt = np.concatenate(uniq_cnt)
            first = np.concatenate(uniq_first)
            order = np.lexsort([first, -cnt])[: self.max_f_grams]
            order = order[cnt[order] >= self.min_freq]
            rows = rows[order]
            if rows.size and (rows.max() > 0x7FFFFFFF or np.any((rows < 0) & (rows != -1))):
                raise ValueError("token ids must be in [0, 2^31)")
            self._tokens = rows.astype(np.int32)
            self._lens = (rows >= 0).sum(axis=1).astype(np.uint8)
        else:
            self._tokens = np.zeros((0, max_n), dtype=np.int32)
            self._lens = np.zeros((0,), dtype=np.uint8)
        self._maps = None
        self._drop_index()
        if verbose:
            print(f"Extracted {len(self)} f-grams")
        return self

    def fit_device(self, tokenized_texts: Iterable[Sequence[int]], verbose: bool = True,
                   device: Optional[torch.device] = None) -> "NGramExtractor":
        """``fit`` on the GPU: same vocabulary and ids as :meth:`fit` / the reference.

        ``scone_fit_vocab`` (``csrc/fit.cu``): every n-gram occurrence is hashed to 64 bits in the reference's enumeration
        order (text, n, start), the hashes are radix-sorted (stable: the first member of a run is the n-gram's first
        occurrence), runs are counted and verified to hold a single n-gram, then ranked by (count descending, first
        occurrence ascending); truncate to ``max_f_grams``, then drop counts below ``min_freq`` (reference :91-94, in
        that order).  No ``[M, n]`` key rows are materialised: 32 bytes of scratch per occurrence.
        """
        import ctypes as C
        from .. import _lib
        from ..index import _stream_ptr
        dev = torch.device(device) if device is not None else (torch.device(self.device) if self.device else _default_device())
        if dev.type != "cuda":
            raise ValueError("scone_b200 has no CPU path: device must be CUDA")
        if dev.index is None:
            dev = torch.device("cuda", torch.cuda.current_device())
        texts = [np.asarray(t, dtype=np.int64).ravel() for t in tokenized_texts]
        max_n = self.max_n
        lens_np = np.array([len(t) for t in texts], dtype=np.int64)
        M = int(lens_np.sum())
        self._tokens = np.zeros((0, max_n), dtype=np.int32)
        self._lens = np.zeros((0,), dtype=np.uint8)
        if M and self.max_f_grams > 0:
            flat64 = np.concatenate(texts)
            if flat64.min() < 0 or flat64.max() > 0x7FFFFFFF:
                raise ValueError("token ids must be in [0, 2^31)")
            flat = torch.from_numpy(flat64.astype(np.int32)).to(dev)
            offs = torch.from_numpy(np.concatenate([[0], np.cumsum(lens_np)]).astype(np.int64)).to(dev)
            cap = int(min(self.max_f_grams, M * max_n))
            out_tok = torch.empty((cap, max_n), dtype=torch.int32, device=dev)
            out_len = torch.empty((cap,), dtype=torch.uint8, device=dev)
            n_out, distinct = C.c_int64(0), C.c_int64(0)
            L = _lib.load()
            with torch.cuda.device(dev):
                for attempt in range(4):               # a 64-bit hash shared by two n-grams is detected, never trusted: new seed
                    rc = L.scone_fit_vocab(flat.data_ptr(), M, offs.data_ptr(), len(texts), max_n, int(self.min_freq), cap,
                                           C.c_uint64(0x5C0E + attempt), out_tok.data_ptr(), out_len.data_ptr(), None,
                                           C.byref(n_out), C.byref(distinct), _stream_ptr(dev))
                    if rc != _lib.E_VOCAB:
                        break
                _lib.check(rc)
            k = int(n_out.value)
            self._tokens = out_tok[:k].cpu().numpy()
            self._lens = out_len[:k].cpu().numpy()
        self._maps = None
        self._drop_index()
        if verbose:
            print(f"Extracted {len(self)} f-grams")
        return self

    # ---- device index -----------------------------------------------------------------------------
    def device_index(self, device: Optional[torch.device] = None, load_factor: float = 0.0) -> FGramIndex:
        """The GPU hash index of the current vocabulary (built once, cached)."""
        dev = torch.device(device) if device is not None else (torch.device(self.device) if self.device else _default_device())
        if dev.type != "cuda":
            raise ValueError("scone_b200 has no CPU path: device must be CUDA")
        if dev.index is None:
            dev = torch.device("cuda", torch.cuda.current_device())
        if self._index is None or self._index.device != dev:
            self._drop_index()
            if self._tokens.shape[1] > self.max_n:
                # the reference only ever tests n <= max_n (n_gram_extractor.py:119), so such f-grams could never match
                raise ValueError(f"the vocabulary holds f-grams longer than max_n = {self.max_n}; raise max_n or drop them")
            toks = torch.from_numpy(self._tokens).to(dev)
            lens = torch.from_numpy(self._lens).to(dev)
            self._index = FGramIndex(toks, lens, load_factor=load_factor)
        return self._index

    def lookup(self, input_ids: torch.Tensor):
        """Batched longest match: (fgram_id int32 [B, L], match_len uint8 [B, L]) for CUDA ``input_ids`` [B, L]."""
        return self.device_index(input_ids.device).lookup(input_ids)

    # ---- reference match primitive (:106-126) --------------------------------------------------------
    def get_token_f_grams(self, token_ids: List[int]) -> Dict[int, List[FGram]]:
        """For each position, all f-grams that contain it (order: ascending n, then ascending start)."""
        token_ids = list(token_ids)
        L = len(token_ids)
        out: Dict[int, List[FGram]] = {i: [] for i in range(L)}
        if L == 0 or len(self) == 0:
            return out
        index = self.device_index()
        ids = torch.tensor([token_ids], dtype=torch.long, device=index.device)
        hit = (index.match_all(ids)[0] >= 0).cpu().numpy()          # [L, max_n], slot n-1 = n-gram ENDING there
        for n in range(1, min(self.max_n + 1, L + 1)):
            ends = np.flatnonzero(hit[:, n - 1])
            for e in ends:                                           # ascending end == ascending start
                i = int(e) - n + 1
                g = tuple(token_ids[i:i + n])
                for j in range(i, i + n):
                    out[j].append(g)
        return out

    # ---- persistence (reference :128-166; same on-disk format) ------------------------------------------
    def save(self, path: str) -> None:
        data = {
            "max_n": self.max_n,
            "min_freq": self.min_freq,
            "max_f_grams": self.max_f_grams,
            "f_gram_to_id": {",".join(map(str, k)): v for k, v in self.f_gram_to_id.items()},
        }
        np.save(path, data, allow_pickle=True)

    @classmethod
    def load(cls, path: str) -> "NGramExtractor":
        data = np.load(path, allow_pickle=True).item()
        ex = cls(max_n=data["max_n"], min_freq=data["min_freq"], max_f_grams=data["max_f_grams"])
        ex.f_gram_to_id = {tuple(map(int, k.split(","))): v for k, v in data["f_gram_to_id"].items()}
        return ex
