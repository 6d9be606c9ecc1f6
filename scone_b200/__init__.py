"""scone_b200 -- B200-native (sm_100a) implementation of SCONE's inference-time input-embedding lookup.

Host code is Python/PyTorch behind the reference's ``NGramExtractor`` / ``EmbeddingCache`` API; the
hot path is hand-written CUDA reached through the C ABI of ``include/scone_b200.h``.  No CPU fallback.
"""

__version__ = "0.2.0"

from .index import FGramIndex  # noqa: F401
from .table import CacheTable, embed_forward, embed_gather, embed_mean_forward, table_layout  # noqa: F401
from .tokenization.n_gram_extractor import NGramExtractor  # noqa: F401
from .inference.embedding_cache import EmbeddingCache  # noqa: F401
from .models.input_embedding import SconeInputEmbedding  # noqa: F401
from .pipeline import HostPipeline  # noqa: F401
