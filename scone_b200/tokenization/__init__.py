"""Mirror of ``scone.tokenization`` (hot-path part)."""
from .n_gram_extractor import NGramExtractor  # noqa: F401
