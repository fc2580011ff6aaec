"""vsearch_b200 -- B200-native engine for vsearch's index-scoring hot path
(``Retriever.retrieve`` -> ``Index.search`` -> top-k).  See DESIGN.md."""
from .index import (BoTIndex, Index, IndexType, SearchResults, SparseIndex, dense_to_csr, merge_keys,  # noqa: F401
                    topk_sparsify)
from .retriever import Retriever  # noqa: F401
from .sharded import ShardedIndex, row_partition  # noqa: F401

__all__ = ["topk_sparsify", "dense_to_csr", "Retriever", "Index", "SparseIndex", "BoTIndex", "IndexType", "SearchResults", "ShardedIndex",
           "row_partition", "merge_keys"]
