"""financial_rag_b200 -- B200-native exact vector search behind Financial-RAG's ChildVectorStore API.

One hot path only: child-chunk similarity scan + top-k + cross-collection
fusion, as hand-written sm_100a CUDA behind the C ABI of include/fr_index.h.  There is no CPU
fallback: importing the compute entry points without libfrb200.so raises ImportError.
"""
from .child_store import B200ChildStore, ChildChunk
from .collection import B200Client, B200Collection, PersistentClient, reset_registry
from .encoder import B200QueryEncoder
from .ensemble import EnsembleSearcher
from .group import ShardGroup
from .index import (ShardIndex, score_fuse_device, score_fuse_host, canonical_space, maxsim_aggregate_device, maxsim_aggregate_host, merge_shards_device,
                    rrf_fuse_device, rrf_fuse_host)
from .multivector_store import B200MultiVectorChildStore
from .vector_store_factory import get_child_vector_store

__all__ = [
    "B200ChildStore", "ChildChunk", "B200Client", "B200Collection", "B200MultiVectorChildStore", "PersistentClient", "B200QueryEncoder", "EnsembleSearcher", "ShardGroup", "ShardIndex", "score_fuse_device", "score_fuse_host",
    "canonical_space", "get_child_vector_store", "maxsim_aggregate_device", "maxsim_aggregate_host",
    "merge_shards_device", "reset_registry", "rrf_fuse_device", "rrf_fuse_host",
]
