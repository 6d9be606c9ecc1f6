"""Input-embedding forward of the SCONE language model, B200-native.

Mirror of the embedding section of the reference's ``SconeLanguageModel.forward``
(``scone/models/language_model.py:234-258``): it produces the ``inputs_embeds`` tensor [B, L, H] that
is handed to ``base_model.transformer(inputs_embeds=...)``.  Per Algorithm 2 (``assets/algorithm.png``)
the f-gram row REPLACES the token embedding where an f-gram ends (the reference code adds a projected
mean instead), and the optional position add (:253-254) is fused into the same kernel.
``combine="add"`` selects the reference code's ``wte(input_ids) + f_gram_embeddings`` (:239-243) instead, with the
row of the longest f-gram ending at the position as the f-gram term; still one kernel, one rounding.
"""

from __future__ import annotations

from typing import Optional

import torch

from ..inference.embedding_cache import EmbeddingCache


class SconeInputEmbedding(torch.nn.Module):
    """``inputs_embeds = lookup(input_ids) (+ wpe[position])`` in one CUDA kernel."""

    def __init__(self, embedding_cache: EmbeddingCache, wte_weight: torch.Tensor, wpe_weight: Optional[torch.Tensor] = None,
                 combine: str = "replace"):
        super().__init__()
        if combine not in ("replace", "add"):
            raise ValueError("combine must be 'replace' or 'add'")
        self.combine = combine
        self.embedding_cache = embedding_cache
        embedding_cache.set_base_embedding(wte_weight, wpe_weight)
        self.has_positions = wpe_weight is not None

    @torch.no_grad()
    def forward(self, input_ids: torch.Tensor, return_match: bool = False):
        embeds, fgram_id, match_len = self.embedding_cache.lookup(input_ids, add_positions=self.has_positions, combine=self.combine)
        return (embeds, fgram_id, match_len) if return_match else embeds
