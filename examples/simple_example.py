"""The lookup part of the reference's examples/simple_example.py (steps 2 and 6-8) on scone_b200, with synthetic tokens and
rows in place of the tokenizer and the f-gram model (both outside this package's scope).  Needs a B200 (no CPU path).

    python examples/simple_example.py
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import scone_b200 as sb  # noqa: E402


def main():
    dev = torch.device("cuda", 0)
    rng = np.random.default_rng(0)
    V, H_f, H, L = 1000, 256, 768, 128

    print("Step 2: extract f-grams (same ids as the reference's NGramExtractor.fit)")
    phrases = [rng.integers(0, V, size=rng.integers(2, 4)).tolist() for _ in range(300)]
    texts = []
    for _ in range(400):                                    # a corpus in which some phrases recur
        t = []
        while len(t) < L:
            t += phrases[rng.integers(0, len(phrases))] if rng.random() < 0.6 else [int(rng.integers(0, V))]
        texts.append(t[:L])
    extractor = sb.NGramExtractor(max_n=3, min_freq=5, max_f_grams=10_000).fit(texts, verbose=False)
    print(f"  {len(extractor.f_grams)} f-grams, e.g. {list(extractor.f_gram_to_id.items())[:3]}")

    print("Step 6-7: precomputed f-gram embeddings -> cache (the f_gram_projection H_f -> H is folded in on the tensor cores)")
    n = len(extractor.f_grams)
    fgram_rows = torch.from_numpy(rng.normal(0, 0.02, size=(n, H_f)).astype(np.float32))   # stands in for the f-gram model
    projection = torch.from_numpy(rng.normal(0, H_f ** -0.5, size=(H, H_f)).astype(np.float32))
    cache = sb.EmbeddingCache(extractor, embedding_dim=H, quant="int8", out_dtype=torch.bfloat16, device=dev)
    cache.cache_embeddings(list(range(n)), fgram_rows, verbose=False, projection=projection)
    wte = torch.from_numpy(rng.normal(0, 0.02, size=(V, H)).astype(np.float32)).to(dev).bfloat16()
    wpe = torch.from_numpy(rng.normal(0, 0.02, size=(L, H)).astype(np.float32)).to(dev).bfloat16()
    cache.set_base_embedding(wte, wpe)

    print("Step 8: the engine's per-position loop is ONE call: longest f-gram ending at each position, row or fallback, + wpe")
    ids = torch.tensor(texts[:8], dtype=torch.long, device=dev)
    embeds, fgram_id, match_len = cache.lookup(ids, add_positions=True)
    torch.cuda.synchronize()
    print(f"  inputs_embeds {tuple(embeds.shape)} {embeds.dtype}; {(fgram_id >= 0).float().mean().item():.0%} of the positions hit an f-gram")
    b, i = map(int, (fgram_id >= 0).nonzero()[0])
    gram = tuple(ids[b, i + 1 - int(match_len[b, i]): i + 1].tolist())
    assert extractor.f_gram_to_id[gram] == int(fgram_id[b, i])
    print(f"  position ({b}, {i}): f-gram {gram} -> id {int(fgram_id[b, i])}")

    print("The reference's own methods still work (global f-gram ids, fp32 rows):")
    rows = cache.get_embeddings([0, 1, 2])
    per_pos = cache.get_token_embeddings(texts[0][:16])
    print(f"  get_embeddings -> {tuple(rows.shape)}; get_token_embeddings -> rows at positions {sorted(per_pos)[:6]} ...")
    print("Example completed successfully!")


if __name__ == "__main__":
    main()
