"""Mirror of ``scone.inference`` (hot-path part)."""
from .embedding_cache import EmbeddingCache  # noqa: F401
