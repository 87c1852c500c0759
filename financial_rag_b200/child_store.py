"""B200ChildStore -- drop-in for the reference's ``ChromaChildStore``.

Mirrors parent_child/chroma_child_store.py:9-80 member for member (constructor arguments, env
fallbacks, ``collection_name`` / ``persist_dir`` / ``client`` / ``col`` attributes, ``upsert_children``,
``search``, ``count``) so rag_backend.py:632-699, parent_child/retriever.py:55-88,
parent_child/pipeline.py:137-143, ingest_all.py:40-41 and check_collections.py:19-21 call it unchanged.
The chromadb client is replaced by ``B200Client`` (collection.py), whose ``query`` runs the exact
scan on the GPU.  Error behaviour follows the reference: ``search`` / ``upsert_children`` propagate
exceptions, ``count`` swallows them and returns -1.
"""
from __future__ import annotations

import os
from dataclasses import dataclass
from typing import Any, Dict, List, Optional

import numpy as np

from .collection import B200Client


@dataclass
class ChildChunk:
    """The record ``upsert_children`` consumes -- same fields as the reference's
    parent_child/parent_child_chunker.py:38-44 (any object with these attributes works)."""
    child_id: int
    parent_id: int
    content: str
    embedding: Optional[List[float]] = None
    context: Optional[str] = None


def _has_embedding(e) -> bool:
    """``if not getattr(c, "embedding", None)`` (chroma_child_store.py:41) without tripping over
    numpy arrays / tensors, which the reference would reject with 'truth value is ambiguous'."""
    if e is None:
        return False
    try:
        return len(e) > 0
    except TypeError:
        return True


def _as_query_vector(text_vector) -> np.ndarray:
    """list[float] (rag_backend.py:699), 1-D/2-D numpy, or a torch tensor of shape (d,) or (1,d)
    (retriever.py:87 with convert_to_numpy=False; local_embedder.py:187-191 returns 2-D)."""
    if hasattr(text_vector, "detach"):
        text_vector = text_vector.detach().cpu().numpy()
    v = np.asarray(text_vector, dtype=np.float32)
    if v.ndim == 2 and v.shape[0] == 1:
        v = v[0]
    if v.ndim != 1:
        raise ValueError(f"search expects one query vector, got shape {v.shape}")
    return v


class B200ChildStore:
    """Child vector store backed by the B200 exact-scan index.

    Env vars (same names as the reference so deployments need no new configuration):
      - CHROMA_CHILD_PERSIST_DIR (default: <project root>/.chroma_children)
      - CHILD_VECTOR_COLLECTION (default: parent_child_children)
    and two of our own: B200_CHILD_DTYPE (bf16 | f32), B200_CHILD_DEVICE (CUDA ordinal).
    """

    def __init__(self, persist_dir: str | None = None, collection: str | None = None):
        if persist_dir:
            self.persist_dir = persist_dir
        else:
            env_dir = os.getenv("CHROMA_CHILD_PERSIST_DIR")
            if env_dir:
                self.persist_dir = env_dir
            else:
                project_root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
                self.persist_dir = os.path.join(project_root, ".chroma_children")
        self.collection_name = collection or os.getenv("CHILD_VECTOR_COLLECTION", "parent_child_children")
        # The registry key is the directory name; the collection persists itself under
        # <persist_dir>/<collection>.b200/ at its first mutation (collection.py).
        self.client = B200Client(path=self.persist_dir)
        self.col = self.client.get_or_create_collection(name=self.collection_name, metadata={"hnsw:space": "cosine"})

    def upsert_children(self, children) -> bool:
        ids: List[str] = []
        metadatas: List[Dict[str, Any]] = []
        embeddings: List[Any] = []
        for c in children:
            emb = getattr(c, "embedding", None)
            if not _has_embedding(emb):
                continue
            ids.append(str(c.child_id))
            # Store full child content as snippet (no truncation), like the reference
            meta = {"parent_id": str(c.parent_id), "snippet": c.content}
            if getattr(c, "context", None):
                meta["context"] = c.context
            if getattr(c, "document_id", None) is not None:  # extension: lets the doc-level helpers below work
                meta["document_id"] = str(c.document_id)
            metadatas.append(meta)
            if hasattr(emb, "detach"):
                emb = emb.detach().cpu().numpy()
            embeddings.append(np.asarray(emb, dtype=np.float32).reshape(-1))
        if not ids:
            return True
        self.col.upsert(ids=ids, embeddings=np.stack(embeddings), metadatas=metadatas)
        return True

    def search(self, text_vector, top_k: int = 6):
        res = self.col.query(query_embeddings=_as_query_vector(text_vector)[None, :], n_results=top_k,
                             include=["metadatas", "distances"])
        out: List[Dict[str, Any]] = []
        ids = res.get("ids", [[]])[0]
        dists = res.get("distances", [[]])[0]
        metas = res.get("metadatas", [[]])[0]
        for i in range(len(ids)):
            meta = metas[i] or {}
            dist = dists[i]
            # Convert distance to score similar to cosine similarity (chroma_child_store.py:72)
            score = 1.0 - float(dist) if dist is not None else None
            out.append({"score": score, "child_id": ids[i], "payload": meta})
        return out

    def search_batch(self, text_vectors, top_k: int = 6):
        """B queries in one GPU pass (the reference issues them one by one, rag_backend.py:675-714)."""
        res = self.col.query(query_embeddings=text_vectors, n_results=top_k, include=["metadatas", "distances"])
        out = []
        for ids, dists, metas in zip(res["ids"], res["distances"], res["metadatas"]):
            out.append([{"score": 1.0 - float(d), "child_id": i, "payload": m or {}}
                        for i, d, m in zip(ids, dists, metas)])
        return out

    def count(self) -> int:
        try:
            return int(self.col.count())
        except Exception:
            return -1

    # The reference probes both with hasattr() and finds neither on ChromaChildStore
    # (api_server.py:230-231, 267-270); children ingested with a ``document_id`` attribute support them.
    def count_for_document(self, doc_id) -> int:
        return len(self.col.keys_where("document_id", str(doc_id)))

    def delete_by_document_id(self, doc_id) -> int:
        keys = self.col.keys_where("document_id", str(doc_id))
        if keys:
            self.col.delete(where={"document_id": str(doc_id)})
        return len(keys)
