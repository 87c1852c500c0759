"""Hybrid child retrieval: batched dense scans + on-GPU fusion, BM25 over the candidates, merge.

Replaces the retrieval core of ``_retrieve_children_hybrid`` (rag_backend.py:653-832): what happens once the query
variants and the encoder ensemble exist (the LLM query normaliser and the embedder construction,
rag_backend.py:578-650, stay with the caller).  The observable behaviour is the reference's -- list order (variants
outer, encoders inner, multi-vector lists first), ``1/(k + rank)`` fusion, BM25 over snippet (+ context) per
variant with the maximum kept, ``dense + bm25 / len(corpus)``, stable descending order, the output records -- but the
work is organised around arrays:

  dense     every encoder answers ALL query variants in one batched scan (``B200Collection.search_arrays``: one
            corpus pass instead of one per (variant, encoder) pair, rag_backend.py:675-714); hits stay
            ``(distance, key)`` matrices until the candidates are numbered;
  fusion    candidates get dense ordinals in first-seen order (the order the reference's dicts iterate in, which
            decides ties); the ``[lists, 1, k]`` ordinal matrix goes through the K5 kernel (``fusion="rrf"``,
            rag_backend.py:720-731) or K5b (``fusion="avg"``, rag_backend.py:732-754) -- fp64, bit-exact;
  lexical   ``Bm25Index``: vocabulary ids, a dense term-frequency matrix and idf / length vectors built once for
            the <= few hundred candidates; a query is a handful of column operations.  ``rank_bm25`` (a
            requirements.txt dependency the reference tree does not vendor) is restated with its published
            defaults (k1 = 1.5, b = 0.75, epsilon = 0.25) and its float64 operation order;
  merge     two vectors and one stable argsort (rag_backend.py:790-798), then the records (rag_backend.py:820-832).
A BM25 failure (e.g. every snippet empty) degrades to dense-only scores, as the reference's try/except does
(rag_backend.py:777-788).
"""
from __future__ import annotations

import math
import os
from typing import Any, Dict, List, Optional, Sequence, Tuple

import numpy as np

from .index import rrf_fuse_host, score_fuse_host


# ---------------------------------------------------------------------------------------------------------------
class Bm25Index:
    """Okapi BM25 over a small tokenised corpus, array form.

    ``tf[d, w]`` is the count of vocabulary word ``w`` in document ``d`` (dense: the corpus is the fused candidate
    set, <= a few hundred snippets), ``idf[w]`` follows rank_bm25 0.2.2 (``ln(N - n + 0.5) - ln(n + 0.5)``, negative
    values replaced by ``epsilon`` times the mean idf, the mean summed in vocabulary order), and a query's score
    vector is accumulated token by token as
    ``idf * (f * (k1 + 1) / (f + k1 * (1 - b + b * len / avg_len)))`` -- the same float64 operations in the same
    order, so the numbers equal rank_bm25's bit for bit."""

    def __init__(self, documents: Sequence[Sequence[str]], k1: float = 1.5, b: float = 0.75, epsilon: float = 0.25):
        self.k1, self.b = float(k1), float(b)
        self.n_docs = len(documents)
        self.word_id: Dict[str, int] = {}
        rows: List[int] = []
        cols: List[int] = []
        for d, doc in enumerate(documents):
            for word in doc:
                rows.append(d)
                cols.append(self.word_id.setdefault(word, len(self.word_id)))
        n_words = len(self.word_id)
        self.tf = np.zeros((self.n_docs, max(n_words, 1)), dtype=np.int64)
        if rows:
            np.add.at(self.tf, (np.asarray(rows), np.asarray(cols)), 1)
        self.doc_len = np.fromiter((len(doc) for doc in documents), dtype=np.int64, count=self.n_docs)
        if self.n_docs == 0 or n_words == 0:
            raise ValueError("BM25 needs at least one non-empty document")
        self.avg_len = int(self.doc_len.sum()) / self.n_docs
        doc_freq = (self.tf > 0).sum(axis=0).tolist()
        idf = [math.log(self.n_docs - n + 0.5) - math.log(n + 0.5) for n in doc_freq]
        mean = 0.0
        for v in idf:  # summed one by one in vocabulary order (not a pairwise numpy sum)
            mean += v
        mean /= n_words
        floor = epsilon * mean
        self.idf = np.array([floor if v < 0 else v for v in idf], dtype=np.float64)
        self._len_term = self.k1 * (1 - self.b + self.b * self.doc_len / self.avg_len)

    def scores(self, query_tokens: Sequence[str]) -> np.ndarray:
        out = np.zeros(self.n_docs)
        for token in query_tokens:
            w = self.word_id.get(token)
            if w is None:
                continue  # unseen word: idf 0, contributes nothing
            f = self.tf[:, w]
            out += self.idf[w] * (f * (self.k1 + 1) / (f + self._len_term))
        return out


class BM25Okapi:
    """``rank_bm25.BM25Okapi``'s two calls as rag_backend.py:779-783 makes them, on top of ``Bm25Index``."""

    def __init__(self, corpus: Sequence[Sequence[str]], k1: float = 1.5, b: float = 0.75, epsilon: float = 0.25):
        self._index = Bm25Index(corpus, k1, b, epsilon)

    def get_scores(self, query: Sequence[str]) -> np.ndarray:
        return self._index.scores(query)


# ---------------------------------------------------------------------------------------------------------------
class _HitList:
    """One ranked list: child ids best first, their scores and payloads (rank = position + 1)."""

    __slots__ = ("ids", "scores", "dists", "payloads")

    def __init__(self, ids, scores, dists, payloads):
        self.ids, self.scores, self.dists, self.payloads = ids, scores, dists, payloads


def _query_matrix(embedder, queries: Sequence[str]) -> np.ndarray:
    """One row per query variant; accepts (d,), (1, d) numpy results and torch tensors (rag_backend.py:678-697)."""
    rows = []
    for q in queries:
        v = embedder.encode(q, convert_to_numpy=True)
        if hasattr(v, "detach"):
            v = v.detach().cpu().numpy()
        v = np.asarray(v, dtype=np.float32)
        rows.append(v.reshape(-1, v.shape[-1])[0])
    return np.stack(rows)


def _dense_lists(store, qmat: np.ndarray, k: int) -> List[_HitList]:
    """The ``k`` nearest children of every row of ``qmat``.  A B200 store answers all rows in one scan and hands back
    matrices; any other store with the reference's ``search`` is asked row by row."""
    col = getattr(store, "col", None)
    if col is not None and hasattr(col, "search_arrays"):
        dist, keys = col.search_arrays(qmat, k)
        out = []
        for drow, krow in zip(dist, keys):
            live = krow != -1
            recs = [col.payload_of_key(int(key)) for key in krow[live]]
            keep = np.array([r is not None for r in recs], dtype=bool)  # a child deleted between the scan and this lookup
            d32 = drow[live][keep]
            out.append(_HitList([r["id"] for r in recs if r is not None],
                                [1.0 - float(d) for d in d32],          # the store's score (chroma_child_store.py:72)
                                d32, [r["metadata"] or {} for r in recs if r is not None]))
        return out
    out = []
    for row in qmat:
        hits = store.search(row.astype(float).tolist(), top_k=k)
        out.append(_HitList([str(h.get("child_id") or "") for h in hits], [float(h.get("score", 0.0) or 0.0) for h in hits],
                            None, [h.get("payload", {}) or {} for h in hits]))
    return out


def _fuse(lists: Sequence[_HitList], fusion: str, k_rrf: int, device: int) -> Tuple[List[str], np.ndarray, List[dict]]:
    """Candidate ids in first-seen order, their fused dense scores, and the payload first seen with each."""
    ordinal: Dict[str, int] = {}
    payload: List[dict] = []
    width = max((len(h.ids) for h in lists), default=0)
    if width == 0:
        return [], np.zeros(0), []
    table = np.full((len(lists), 1, width), -1, dtype=np.int64)
    for l, h in enumerate(lists):
        for r, cid in enumerate(h.ids):
            if not cid:
                continue  # the reference skips hits without a child id
            o = ordinal.get(cid)
            if o is None:
                o = ordinal[cid] = len(payload)
                payload.append(h.payloads[r])
            table[l, 0, r] = o
    names = list(ordinal)
    n = len(names)
    if fusion == "rrf":
        score, order = rrf_fuse_host(table, k_rrf, n, device=device)
    elif all(h.dists is not None and len(h.dists) == len(h.ids) for h in lists):
        dist = np.zeros(table.shape, dtype=np.float32)
        for l, h in enumerate(lists):
            dist[l, 0, :len(h.ids)] = h.dists
        score, order = score_fuse_host(dist, table, n, device=device)
    else:
        # lists whose scores are not ``1 - distance`` of a scan (multi-vector MaxSim sums, foreign stores): the
        # same min-max / mean arithmetic on the host, element by element in list order
        fused = np.zeros(n)
        for h in lists:
            if not h.scores:
                continue
            lo, hi = min(h.scores), max(h.scores)
            for cid, s in zip(h.ids, h.scores):
                if cid:
                    fused[ordinal[cid]] += (s - lo) / (hi - lo) if hi > lo else 0.0
        return names, fused / float(len(lists)), payload
    fused = np.zeros(n)
    live = order[0] != -1
    fused[order[0][live]] = score[0][live]
    return names, fused, payload


def _candidate_text(payload: dict) -> str:
    snippet = payload.get("snippet") or ""
    extra = payload.get("context") or ""
    return (snippet + "\n" + extra).strip() if extra else snippet


def retrieve_children_hybrid(queries: Sequence[str], ensemble: Sequence[Dict[str, Any]], max_children: int = 24, *,
                             fusion: str = "rrf", k_rrf: Optional[int] = None, multivector=None,
                             device: int = 0) -> Tuple[List[Dict[str, Any]], Dict[str, int], List[str]]:
    """``ensemble``: [{"name", "embedder" (``.encode(text, convert_to_numpy=True)``), "vec" (a child store)}].
    Returns (child_chunks, child -> parent map, queries) as rag_backend.py:832 does."""
    queries = list(queries)
    per_encoder = [_dense_lists(m["vec"], _query_matrix(m["embedder"], queries), max_children) for m in ensemble]
    lists: List[_HitList] = []
    if multivector is not None:  # CHILD_USE_MULTIVECTOR=true: its lists come first (rag_backend.py:655-672)
        for q in queries:
            hits = multivector.search_aggregate(q, top_k_children=max_children)
            lists.append(_HitList([str(h.get("child_id") or "") for h in hits],
                                  [float(h.get("score", 0.0) or 0.0) for h in hits], None,
                                  [h.get("payload", {}) or {} for h in hits]))
    for v in range(len(queries)):      # variants outer, encoders inner (rag_backend.py:675-676)
        lists.extend(enc[v] for enc in per_encoder)
    if not lists:
        raise RuntimeError("No child hits from dual-encoder retrieval. Ensure ingestion populated per-model "
                           "collections children_baai_bge_small_en_v1_5 and children_thenlper_gte_small.")
    if k_rrf is None:
        k_rrf = int(os.getenv("ENSEMBLE_RRF_K", "60"))
    names, dense, payloads = _fuse(lists, "rrf" if fusion == "rrf" else "avg", int(k_rrf), device)

    # lexical stage over the candidates that carry text
    texts = [_candidate_text(p) for p in payloads]
    with_text = [i for i, t in enumerate(texts) if t]
    lexical = np.zeros(len(names))
    if with_text:
        try:
            bm25 = Bm25Index([texts[i].split() for i in with_text])
            best = np.zeros(len(with_text))
            for q in queries:
                best = np.maximum(best, bm25.scores(q.split()))
            lexical[with_text] = best
        except Exception:  # noqa: BLE001 - dense-only, like the reference (rag_backend.py:777-788)
            lexical[:] = 0.0
    merged = dense + lexical / (len(with_text) or 1)
    top = np.argsort(-merged, kind="stable")[:max_children]

    child_parent: Dict[str, int] = {}
    for cid, p in zip(names, payloads):
        try:
            if p.get("parent_id") is not None:
                child_parent[cid] = int(p["parent_id"])
        except (TypeError, ValueError):
            pass
    chunks = []
    for i in top.tolist():
        text = texts[i]
        chunks.append({"chunk_id": f"child_{names[i]}", "chunk_text": text, "text": text,
                       "retrieval_score": float(merged[i]), "retrieval_method": "child_hybrid", "child_id": names[i]})
    return chunks, child_parent, queries


__all__ = ["Bm25Index", "BM25Okapi", "retrieve_children_hybrid"]
