"""B200MultiVectorChildStore -- drop-in for the reference's ``MultiVectorChildStore`` (ColBERT-style
late interaction, parent_child/multivector_store.py:27-187) on the B200 exact-scan index.

Same design as the reference: every token embedding of a child chunk is one point with id
``f"{child_id}:{token_idx}"`` and payload ``{child_id, parent_id, token_idx, snippet}``
(multivector_store.py:122-134); a query is embedded per token, each token retrieves its
``topk_per_token`` nearest token vectors, children are scored by MaxSim per token summed over tokens
(multivector_store.py:150-176).  What changes is where the work happens:

  * the reference loops over query tokens and issues one single-vector ``col.query`` each
    (multivector_store.py:150-151); here the T token vectors are ONE batched scan (B = T), and
  * the per-child max / sum / sort runs in the K6 kernel on the hit lists while they are still in HBM
    (``fr_maxsim_aggregate_device``); only the ``top_k_children`` (score, child) pairs come back.

Row keys are ``(child_ordinal << 16) | token_idx`` so the kernel recovers a row's child with a shift;
ordinals are dense per collection and live IN the collection (``B200Collection.group_ordinal``: allocated under
its lock, persisted with it), because the reference constructs these stores per request and side by side
(rag_backend.py:656, pipeline.py:25) and every object must see the same child <-> ordinal map.

Token embedding itself (a BERT forward pass, multivector_store.py:88-111) is SURVEY.md row 8f-4 and not
part of this path: pass ``token_embedder(text, max_tokens) -> list[list[float]]``.  Without one the store
is soft-disabled exactly like the reference without a model (multivector_store.py:80-86): upserts return
True without storing, searches return [].
Env vars are the reference's: CHILD_MULTI_COLLECTION, CHROMA_CHILD_PERSIST_DIR, MULTIVECTOR_MAX_TOKENS,
MULTIVECTOR_QUERY_TOKENS, MULTIVECTOR_TOPK_PER_TOKEN.
"""
from __future__ import annotations

import os
from typing import Any, Callable, Dict, List, Optional

import numpy as np

from .collection import B200Client

TOKEN_BITS = 16  # token_idx < 65536; MULTIVECTOR_MAX_TOKENS defaults to 128


class B200MultiVectorChildStore:
    def __init__(self, persist_dir: str | None = None, collection: str | None = None,
                 token_embedder: Optional[Callable[[str, int], List[List[float]]]] = None):
        self.persist_dir = persist_dir or os.getenv("CHROMA_CHILD_PERSIST_DIR",
                                                    os.path.join(os.getcwd(), ".chroma_children"))
        self.collection_name = collection or os.getenv("CHILD_MULTI_COLLECTION", "parent_child_child_tokens")
        self.client = B200Client(path=self.persist_dir)
        self.col = self.client.get_or_create_collection(name=self.collection_name, metadata={"hnsw:space": "cosine"})
        self.child_max_tokens = int(os.getenv("MULTIVECTOR_MAX_TOKENS", "128"))
        self.query_max_tokens = int(os.getenv("MULTIVECTOR_QUERY_TOKENS", "16"))
        self.topk_per_token = int(os.getenv("MULTIVECTOR_TOPK_PER_TOKEN", "10"))
        self._token_embedder = token_embedder
        self._disabled_reason: Optional[str] = None if token_embedder else "No suitable model available for multi-vector store"

    # -- helpers ------------------------------------------------------------------------------------
    def _ordinal(self, child_id: str) -> int:
        """child id -> dense ordinal (the upper bits of the row keys); the map belongs to the collection."""
        return self.col.group_ordinal(child_id)

    def _embed_tokens(self, text: str, max_tokens: int) -> List[List[float]]:
        if not text or self._disabled_reason:
            return []
        vecs = np.asarray(self._token_embedder(text, max_tokens), dtype=np.float32)
        if vecs.size == 0:
            return []
        # L2-normalise token vectors like the reference (multivector_store.py:109-110)
        vecs = vecs / np.maximum(np.linalg.norm(vecs, axis=1, keepdims=True), 1e-12)
        return vecs

    # -- the reference's surface ----------------------------------------------------------------------
    def upsert_child_tokens(self, children: List[Any]) -> bool:
        ids: List[str] = []
        keys: List[int] = []
        metas: List[Dict[str, Any]] = []
        embs: List[np.ndarray] = []
        for c in children:
            text = getattr(c, "content", None)
            if not text:
                continue
            token_vecs = self._embed_tokens(text, self.child_max_tokens)
            if len(token_vecs) == 0:
                continue
            child_id = str(getattr(c, "child_id"))
            parent_id = str(getattr(c, "parent_id"))
            base = self._ordinal(child_id) << TOKEN_BITS
            for idx, v in enumerate(token_vecs[: 1 << TOKEN_BITS]):
                ids.append(f"{child_id}:{idx}")
                keys.append(base | idx)
                metas.append({"child_id": child_id, "parent_id": parent_id, "token_idx": idx, "snippet": text})
                embs.append(v)
        if not ids:
            return True
        self.col.upsert(ids=ids, embeddings=np.stack(embs), metadatas=metas, keys=keys)
        return True

    def search_aggregate(self, query_text: str, top_k_children: int = 24) -> List[Dict[str, Any]]:
        qvecs = self._embed_tokens(query_text, self.query_max_tokens)
        if len(qvecs) == 0:
            return []
        return self.search_aggregate_vectors(qvecs, top_k_children)

    def search_aggregate_vectors(self, qvecs, top_k_children: int = 24) -> List[Dict[str, Any]]:
        """The aggregation for already-embedded query tokens ([T, dim])."""
        import torch

        from .index import maxsim_aggregate_device

        ix = self.col.index
        if ix is None or ix.count() == 0:
            return []
        from .group import ShardGroup

        q_host = torch.as_tensor(np.ascontiguousarray(qvecs, dtype=np.float32))
        t, kp = q_host.shape[0], self.topk_per_token
        if isinstance(ix, ShardGroup):                            # row-sharded collection: the merged lists land on its first GPU
            qs = [q_host.to(torch.device("cuda", d)) for d in ix.devices]
            dl, kl = ix.search_device(qs, kp, merge_on=[0])
            dist, keys = dl[0], kl[0]
        else:
            dist, keys = ix.search_device(q_host.to(torch.device("cuda", ix.device)), kp)   # ONE scan for all T tokens
        sc, grp = maxsim_aggregate_device(dist.view(1, t, kp), keys.view(1, t, kp), TOKEN_BITS, top_k_children)
        sc, grp = sc[0].cpu().tolist(), grp[0].cpu().tolist()
        out: List[Dict[str, Any]] = []
        for s, g in zip(sc, grp):
            if g == -1:
                break
            cid = self.col.group_name(g)
            if cid is None:
                continue
            meta = self.col.metadata_of_key(g << TOKEN_BITS) or {}
            out.append({"score": float(s), "child_id": cid,
                        "payload": {"parent_id": meta.get("parent_id"), "snippet": meta.get("snippet", "")}})
        return out
