"""Vocabulary construction on the device at scale (development tool): scone_fit_vocab on a synthetic corpus.

    python tools/bench_fit.py [tokens] [max_n] [max_f_grams]
"""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import scone_b200 as sb  # noqa: E402

M = int(float(sys.argv[1])) if len(sys.argv) > 1 else 100_000_000
max_n = int(sys.argv[2]) if len(sys.argv) > 2 else 4
cap = int(float(sys.argv[3])) if len(sys.argv) > 3 else 10_000_000
V, text_len = 50_257, 1000
rng = np.random.default_rng(0)
u = rng.random(M)
flat = np.clip(np.floor(np.exp(u * np.log(V))).astype(np.int64) - 1, 0, V - 1)      # Zipf-like token stream
del u
texts = [flat[i:i + text_len] for i in range(0, M, text_len)]
torch.cuda.init()
torch.zeros(1, device="cuda")
ex = sb.NGramExtractor(max_n=max_n, min_freq=2, max_f_grams=cap)
t0 = time.perf_counter()
ex.fit_device(texts, verbose=False)
torch.cuda.synchronize()
dt = time.perf_counter() - t0
toks, lens = ex.vocab_arrays()
# spot-check against a host count of the corpus' most frequent n-gram of each length
ok = True
for n in range(1, max_n + 1):
    ids = np.flatnonzero(lens == n)
    if len(ids) == 0:
        continue
    g = toks[ids[0], :n]                                   # the most frequent n-gram of this length (lowest id)
    w = np.lib.stride_tricks.sliding_window_view(flat[: min(M, 20_000_000)], n)
    # (texts are 1000 tokens long and n-grams never cross texts: exclude windows that straddle a boundary)
    starts = np.arange(w.shape[0])
    inside = (starts % text_len) + n <= text_len
    cnt = int((np.all(w == g, axis=1) & inside).sum())
    ok = ok and cnt > 0
print(json.dumps({"tokens": M, "texts": len(texts), "max_n": max_n, "max_f_grams": cap, "f_grams_found": len(ex), "seconds_total": dt,
                  "lens_hist": np.bincount(lens, minlength=max_n + 1).tolist(), "spot_check_ok": bool(ok),
                  "peak_device_GB": torch.cuda.max_memory_allocated() / 1e9,
                  "note": "seconds_total includes flattening the corpus on the host and the H2D copy; scratch comes from cudaMallocAsync"}), flush=True)
