"""Hybrid child retrieval: dual-encoder dense search + RRF on the GPU, BM25 over the candidates, merge.

Mirror of the retrieval core of ``_retrieve_children_hybrid`` (rag_backend.py:653-832), i.e. everything
after the query variants and the encoder ensemble exist (the LLM query normaliser and the embedder
construction, rag_backend.py:578-650, stay with the caller).  Same data flow, same arithmetic:

  1. every (query variant x encoder) pair searches its per-encoder collection for ``max_children`` hits
     (rag_backend.py:675-714) -- here ONE batched scan per encoder (B = number of variants) instead of
     one single-vector query per pair;
  2. RRF over all ranked lists, ``1.0 / (k_rrf + rank)`` summed in list order (rag_backend.py:720-731) --
     the K5 kernel, fp64, bit-exact; ``fusion="avg"`` (rag_backend.py:732-754, dead in the reference
     because ``fusion = "rrf"`` is hard-coded at :589) is kept as a host option;
  3. BM25Okapi over the candidate snippets (+ context), per variant, max over variants
     (rag_backend.py:756-788).  ``rank_bm25`` (requirements.txt) is not vendored in the reference tree;
     ``BM25Okapi`` below restates rank_bm25 0.2.2's published algorithm with its defaults
     (k1 = 1.5, b = 0.75, epsilon = 0.25) and its numpy float64 operation order;
  4. ``score = dense + bm25 / len(corpus)``, stable sort descending, cut (rag_backend.py:790-798);
  5. the reference's output records (rag_backend.py:820-832).
"""
from __future__ import annotations

import math
import os
from typing import Any, Dict, List, Optional, Sequence, Tuple

import numpy as np

from .index import rrf_fuse_host


class BM25Okapi:
    """rank_bm25.BM25Okapi as called at rag_backend.py:779-783 (whitespace-tokenised documents)."""

    def __init__(self, corpus: Sequence[Sequence[str]], k1: float = 1.5, b: float = 0.75, epsilon: float = 0.25):
        self.k1, self.b, self.epsilon = k1, b, epsilon
        self.corpus_size = 0
        self.doc_freqs: List[Dict[str, int]] = []
        self.doc_len: List[int] = []
        self.idf: Dict[str, float] = {}
        nd: Dict[str, int] = {}
        num_doc = 0
        for document in corpus:
            self.doc_len.append(len(document))
            num_doc += len(document)
            frequencies: Dict[str, int] = {}
            for word in document:
                frequencies[word] = frequencies.get(word, 0) + 1
            self.doc_freqs.append(frequencies)
            for word in frequencies:
                nd[word] = nd.get(word, 0) + 1
            self.corpus_size += 1
        self.avgdl = num_doc / self.corpus_size
        # idf with the epsilon floor for terms in more than half of the documents
        idf_sum = 0.0
        negative = []
        for word, freq in nd.items():
            idf = math.log(self.corpus_size - freq + 0.5) - math.log(freq + 0.5)
            self.idf[word] = idf
            idf_sum += idf
            if idf < 0:
                negative.append(word)
        self.average_idf = idf_sum / len(self.idf)
        eps = self.epsilon * self.average_idf
        for word in negative:
            self.idf[word] = eps

    def get_scores(self, query: Sequence[str]) -> np.ndarray:
        score = np.zeros(self.corpus_size)
        doc_len = np.array(self.doc_len)
        for q in query:
            q_freq = np.array([(doc.get(q) or 0) for doc in self.doc_freqs])
            score += (self.idf.get(q) or 0) * (q_freq * (self.k1 + 1) /
                                               (q_freq + self.k1 * (1 - self.b + self.b * doc_len / self.avgdl)))
        return score


def _fuse_rrf_gpu(ranked_lists: List[List[Dict[str, Any]]], k_rrf: int, device: int) -> Dict[str, float]:
    """combined_dense of rag_backend.py:720-731 through the K5 kernel.  Child ids become dense ordinals in
    first-seen order (the dict insertion order the reference's result inherits)."""
    ordinal: Dict[str, int] = {}
    kp = max((len(lst) for lst in ranked_lists), default=0)
    if kp == 0:
        return {}
    keys = np.full((len(ranked_lists), 1, kp), -1, dtype=np.int64)
    for l, lst in enumerate(ranked_lists):
        for r in lst:
            cid = str(r.get("child_id") or "")
            if not cid:
                continue
            rank = int(r.get("rank", 1))  # 1-based position the caller stamped on the hit
            keys[l, 0, rank - 1] = ordinal.setdefault(cid, len(ordinal))
    sc, fused = rrf_fuse_host(keys, k_rrf, len(ordinal), device=device)
    names = list(ordinal)
    score_of = {names[int(o)]: float(s) for s, o in zip(sc[0], fused[0]) if o != -1}
    return {cid: score_of[cid] for cid in names}  # first-seen order, like the reference's dict


def _fuse_avg(ranked_lists: List[List[Dict[str, Any]]]) -> Dict[str, float]:
    combined: Dict[str, float] = {}
    for lst in ranked_lists:
        scores = [float(x.get("score", 0.0) or 0.0) for x in lst]
        if not scores:
            continue
        mn, mx = min(scores), max(scores)
        for x, s in zip(lst, scores):
            cid = str(x.get("child_id") or "")
            if not cid:
                continue
            norm = (s - mn) / (mx - mn) if mx > mn else 0.0
            combined[cid] = combined.get(cid, 0.0) + norm
    nlists = float(len(ranked_lists))
    if nlists > 0:
        for cid in list(combined.keys()):
            combined[cid] /= nlists
    return combined


def retrieve_children_hybrid(queries: Sequence[str], ensemble: Sequence[Dict[str, Any]], max_children: int = 24, *,
                             fusion: str = "rrf", k_rrf: Optional[int] = None, multivector=None,
                             device: int = 0) -> Tuple[List[Dict[str, Any]], Dict[str, int], List[str]]:
    """``ensemble``: [{"name", "embedder" (``.encode(text, convert_to_numpy=True)``), "vec" (a child store)}].
    Returns (child_chunks, child->parent map, queries) exactly as rag_backend.py:832 does."""
    queries = list(queries)
    per_member: List[List[List[Dict[str, Any]]]] = []
    for member in ensemble:
        vecs = []
        for q in queries:
            qv = np.asarray(member["embedder"].encode(q, convert_to_numpy=True), dtype=np.float32)
            vecs.append(qv[0] if qv.ndim == 2 else qv)
        vec = member["vec"]
        if hasattr(vec, "search_batch"):
            per_member.append(vec.search_batch(np.stack(vecs), top_k=max_children))  # one scan, B = len(queries)
        else:
            per_member.append([vec.search(v.astype(float).tolist(), top_k=max_children) for v in vecs])

    ranked_lists: List[List[Dict[str, Any]]] = []
    candidate_payloads: Dict[str, Dict[str, Any]] = {}

    def take(res, q, encoder):
        for rank_idx, r in enumerate(res):
            r["query"], r["encoder"], r["rank"] = q, encoder, rank_idx + 1
        ranked_lists.append(res)
        for r in res:
            cid = str(r.get("child_id") or "")
            if cid and cid not in candidate_payloads:
                candidate_payloads[cid] = r

    if multivector is not None:  # CHILD_USE_MULTIVECTOR=true (rag_backend.py:655-672)
        for q in queries:
            take(multivector.search_aggregate(q, top_k_children=max_children), q, "multivector")
    for qi, q in enumerate(queries):  # list order of the reference: variants outer, encoders inner
        for mi, member in enumerate(ensemble):
            take(per_member[mi][qi], q, member["name"])
    if not ranked_lists:
        raise RuntimeError("No child hits from dual-encoder retrieval. Ensure ingestion populated per-model "
                           "collections children_baai_bge_small_en_v1_5 and children_thenlper_gte_small.")

    if fusion == "rrf":
        k = int(os.getenv("ENSEMBLE_RRF_K", "60")) if k_rrf is None else int(k_rrf)
        combined_dense = _fuse_rrf_gpu(ranked_lists, k, device)
    else:
        combined_dense = _fuse_avg(ranked_lists)

    child_docs: Dict[str, str] = {}
    child_parent: Dict[str, int] = {}
    for cid, rhit in candidate_payloads.items():
        payload = rhit.get("payload", {}) or {}
        snippet = payload.get("snippet") or ""
        ctx_extra = payload.get("context") or ""
        text_for_bm25 = (snippet + "\n" + ctx_extra).strip() if ctx_extra else snippet
        if text_for_bm25 and cid not in child_docs:
            child_docs[cid] = text_for_bm25
        try:
            pid = int(payload.get("parent_id")) if payload.get("parent_id") is not None else None
            if pid is not None:
                child_parent[cid] = pid
        except Exception:
            pass

    corpus_ids = list(child_docs.keys())
    corpus_texts = [child_docs[cid] for cid in corpus_ids]
    bm25_scores: Dict[str, float] = {}
    if corpus_texts:
        bm25 = BM25Okapi([txt.split() for txt in corpus_texts])
        for q in queries:
            scores = bm25.get_scores(q.split())
            for idx, s in enumerate(scores):
                cid = corpus_ids[idx]
                bm25_scores[cid] = max(bm25_scores.get(cid, 0.0), float(s))

    child_score_map: Dict[str, float] = {}
    for cid, dscore in combined_dense.items():
        child_score_map[cid] = dscore + bm25_scores.get(cid, 0.0) / (len(corpus_texts) or 1)
    ranked = sorted(child_score_map.items(), key=lambda it: it[1], reverse=True)[:max_children]

    child_chunks = []
    for cid, score in ranked:
        snippet = child_docs.get(cid, "")
        child_chunks.append({"chunk_id": f"child_{cid}", "chunk_text": snippet, "text": snippet,
                             "retrieval_score": float(score), "retrieval_method": "child_hybrid", "child_id": cid})
    return child_chunks, child_parent, queries
