"""Mirror of ``scone.models`` (input-embedding part of the language model)."""
from .input_embedding import SconeInputEmbedding  # noqa: F401
