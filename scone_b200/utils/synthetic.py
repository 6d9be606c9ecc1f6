"""Synthetic vocabularies, token streams and tables of the BASELINE.json shapes (no tokenizer, no network).

Shared by the tests and bench.py.  numpy generators for sizes the Python oracle can hold; torch
device generators (same construction, counter-style so nothing large is staged on the host) for the
10^7..10^8-f-gram configs.  SURVEY.md section 8d describes the recipe: f-grams of length min_n..max_n over
a Zipf-like token distribution, a stream in which vocabulary f-grams are planted with probability
p_plant (ids uniform over the table = worst case for caches), tables ~ N(0, 0.02^2).
"""

from __future__ import annotations

from typing import Optional, Tuple

import numpy as np
import torch


# ---- token distribution: log-uniform ranks (density ~ 1/rank, i.e. Zipf with exponent 1) ---------------------

def zipf_like_numpy(rng: np.random.Generator, size, V: int) -> np.ndarray:
    u = rng.random(size)
    return np.clip(np.floor(np.exp(u * np.log(V))).astype(np.int64) - 1, 0, V - 1)


def zipf_like_torch(gen: torch.Generator, size, V: int, device) -> torch.Tensor:
    u = torch.rand(size, generator=gen, device=device, dtype=torch.float64)
    return (torch.exp(u * float(np.log(V))).floor().to(torch.int64) - 1).clamp_(0, V - 1)


# ---- vocabularies ------------------------------------------------------------------------------------------------

def make_vocab_numpy(N: int, max_n: int, V: int, seed: int = 0, min_n: int = 2, nested: bool = True) -> Tuple[np.ndarray, np.ndarray]:
    """N DISTINCT f-grams (tokens int32 [N, max_n] padded -1, lens uint8 [N]); id = generation order.

    ``nested`` also inserts every prefix (length >= min_n) and sometimes the (n-1)-suffix of a generated
    f-gram, so that planted f-grams hit at most of their positions and the longest-match rule is exercised.
    Python-set de-duplication: use for N up to ~10^6.
    """
    rng = np.random.default_rng(seed)
    min_n = max(1, min(min_n, max_n))
    capacity = sum(float(V) ** n for n in range(min_n, max_n + 1))
    if N > 0.25 * capacity:
        raise ValueError(f"cannot draw {N} distinct f-grams of length {min_n}..{max_n} over {V} tokens")
    seen, grams = set(), []
    while len(grams) < N:
        m = max(1024, N - len(grams))
        lens = rng.integers(min_n, max_n + 1, size=m)
        toks = zipf_like_numpy(rng, (m, max_n), V)
        for i in range(m):
            g = tuple(int(t) for t in toks[i, :lens[i]])
            cands = [g]
            if nested:
                cands += [g[:k] for k in range(len(g) - 1, min_n - 1, -1)]
                if len(g) - 1 >= min_n and (i & 3) == 0:
                    cands.append(g[1:])
            for c in cands:
                if c not in seen and len(grams) < N:
                    seen.add(c)
                    grams.append(c)
    order = rng.permutation(N)                      # ids uncorrelated with family membership
    out = np.full((N, max_n), -1, dtype=np.int32)
    ln = np.zeros((N,), dtype=np.uint8)
    for i, g in enumerate(grams):
        out[order[i], :len(g)] = g
        ln[order[i]] = len(g)
    return out, ln


def make_vocab_device(N: int, max_n: int, V: int, seed: int = 0, device="cuda", return_longest: bool = False):
    """N distinct f-grams generated on the device (tokens int32 [N, max_n] pad -1, lens uint8 [N]).

    Families: one Zipf-like token sequence of length max_n contributes its prefixes of length 2..max_n
    (so a planted f-gram hits at every position but its first, like text covered by frequent n-grams).
    Distinct BY CONSTRUCTION: the first two tokens of a family encode the family number through a bijection
    of [0, V^2).  Ids are an affine permutation of the generation order, so the rows touched by consecutive
    positions are scattered over the whole table.  ``return_longest`` also returns the ids of the
    length-max_n members (the ones to plant for the highest hit rate).
    """
    if max_n < 2:
        raise ValueError("device vocabulary needs max_n >= 2")
    per = max_n - 1
    F = (N + per - 1) // per
    if F > V * V:
        raise ValueError("too many families for V^2 distinct token pairs")
    gen = torch.Generator(device=device)
    gen.manual_seed(seed)
    fam = zipf_like_torch(gen, (F, max_n), V, device).to(torch.int32)
    M = V * V
    a = 0x9E3779B1 % M
    while np.gcd(a, M) != 1:
        a += 1
    f = torch.arange(F, device=device, dtype=torch.int64)
    x = (f * a) % M                                   # a < 2^32, f < 2^31: no overflow
    fam[:, 0] = (x // V).to(torch.int32)
    fam[:, 1] = (x % V).to(torch.int32)
    i = torch.arange(N, device=device, dtype=torch.int64)
    b = 0x85EBCA6B % max(N, 1)
    while np.gcd(b, max(N, 1)) != 1:
        b += 1
    ids = (i * b + 12345) % N                         # affine bijection of [0, N)
    lens_gen = (2 + (i % per)).to(torch.int32)
    toks_gen = fam[i // per]
    k = torch.arange(max_n, device=device, dtype=torch.int32)[None, :]
    toks_gen = torch.where(k < lens_gen[:, None], toks_gen, torch.full_like(toks_gen, -1))
    toks = torch.empty_like(toks_gen)
    lens = torch.empty((N,), dtype=torch.uint8, device=device)
    toks[ids] = toks_gen
    lens[ids] = lens_gen.to(torch.uint8)
    if return_longest:
        return toks.contiguous(), lens, ids[lens_gen == max_n].contiguous()
    return toks.contiguous(), lens


# ---- token streams -----------------------------------------------------------------------------------------------------

def make_stream_numpy(vocab_tokens: np.ndarray, vocab_lens: np.ndarray, B: int, L: int, V: int, seed: int = 1,
                      p_plant: float = 0.8, pick_ids: Optional[np.ndarray] = None) -> np.ndarray:
    """[B, L] int64 stream of pieces: with probability p_plant a whole vocabulary f-gram (id uniform over
    ``pick_ids`` or the whole vocabulary), else one Zipf-like token.  Pieces may straddle row ends."""
    rng = np.random.default_rng(seed)
    T, N, max_n = B * L, len(vocab_lens), vocab_tokens.shape[1]
    out = zipf_like_numpy(rng, T + max_n, V)
    if N:
        planted = rng.random(T) < p_plant
        pick = rng.integers(0, N, size=T) if pick_ids is None else np.asarray(pick_ids)[rng.integers(0, len(pick_ids), size=T)]
        plen = np.where(planted, vocab_lens[pick].astype(np.int64), 1)
        off = np.cumsum(plen) - plen
        keep = off < T
        for k in range(max_n):
            m = keep & planted & (k < plen)
            out[off[m] + k] = vocab_tokens[pick[m], k]
    return out[:T].reshape(B, L).astype(np.int64)


def make_stream_device(vocab_tokens: torch.Tensor, vocab_lens: torch.Tensor, B: int, L: int, V: int, seed: int = 1,
                       p_plant: float = 0.8, pick_ids: Optional[torch.Tensor] = None) -> torch.Tensor:
    device = vocab_tokens.device
    gen = torch.Generator(device=device)
    gen.manual_seed(seed)
    T, N, max_n = B * L, vocab_lens.numel(), vocab_tokens.shape[1]
    out = zipf_like_torch(gen, (T + max_n,), V, device)
    if N:
        planted = torch.rand((T,), generator=gen, device=device) < p_plant
        pick = torch.randint(0, N, (T,), generator=gen, device=device)
        if pick_ids is not None:
            pick = pick_ids[torch.randint(0, pick_ids.numel(), (T,), generator=gen, device=device)]
        plen = torch.where(planted, vocab_lens[pick].to(torch.int64), torch.ones_like(pick))
        off = torch.cumsum(plen, 0) - plen
        keep = off < T
        for k in range(max_n):
            m = keep & planted & (k < plen)
            out[off[m] + k] = vocab_tokens[pick[m], k].to(torch.int64)
    return out[:T].reshape(B, L).contiguous()


# ---- tables ----------------------------------------------------------------------------------------------------------------

def make_rows_numpy(N: int, D: int, seed: int = 2, std: float = 0.02) -> np.ndarray:
    return (np.random.default_rng(seed).standard_normal((N, D)) * std).astype(np.float32)


def pack_table_numpy(quant: str, payload: np.ndarray, scales: Optional[np.ndarray], align: int = 32):
    """Lay an oracle table out in the product's row format: payload, then scale(s), stride rounded to ``align``.
    Returns (packed uint8 [N, row_stride], row_stride, scale_offset)."""
    N = payload.shape[0]
    pb = np.ascontiguousarray(payload).view(np.uint8).reshape(N, -1)
    if quant in ("fp16", "fp32"):
        sb = np.zeros((N, 0), dtype=np.uint8)
    elif quant == "int8":
        sb = np.ascontiguousarray(scales, dtype=np.float32).reshape(N, 1).view(np.uint8)
    else:
        sb = np.ascontiguousarray(scales, dtype=np.float16).view(np.uint8).reshape(N, -1)
    scale_off = pb.shape[1] if quant in ("int8", "int4") else 0
    used = pb.shape[1] + sb.shape[1]
    stride = (used + align - 1) // align * align
    out = np.zeros((N, stride), dtype=np.uint8)
    out[:, :pb.shape[1]] = pb
    out[:, pb.shape[1]:used] = sb
    return out, stride, scale_off


def fill_table_device(table, seed: int = 2, std: float = 0.02, chunk_bytes: int = 1 << 30) -> None:
    """Fill a CacheTable with quantised N(0, std^2) rows, generated and quantised on the device in chunks."""
    gen = torch.Generator(device=table.device)
    gen.manual_seed(seed)
    chunk = max(1, chunk_bytes // (4 * table.dim))
    for s in range(0, table.num_rows, chunk):
        k = min(chunk, table.num_rows - s)
        rows = torch.randn((k, table.dim), generator=gen, device=table.device, dtype=torch.float32) * std
        table.store(rows, row_base=s)


def make_base_device(V: int, D: int, dtype=torch.bfloat16, seed: int = 3, std: float = 0.02, device="cuda") -> torch.Tensor:
    gen = torch.Generator(device=device)
    gen.manual_seed(seed)
    return (torch.randn((V, D), generator=gen, device=device, dtype=torch.float32) * std).to(dtype).contiguous()
