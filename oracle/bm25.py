"""CPU ORACLE (test infrastructure, NOT product code) -- BM25 + dense merge of the hybrid retriever.

Restates rag_backend.py:756-798 in plain Python floats:
  * ``rank_bm25.BM25Okapi`` as used at rag_backend.py:779-783.  The module is a requirements.txt dependency
    (``rank-bm25``, unpinned) that is NOT vendored in /root/reference and not installed here, so this is
    its published algorithm (rank_bm25 0.2.2: k1 = 1.5, b = 0.75, epsilon = 0.25;
    idf = ln(N - n + 0.5) - ln(n + 0.5), negative idfs replaced by epsilon * mean idf;
    score = sum_q idf(q) * f (k1 + 1) / (f + k1 (1 - b + b dl / avgdl))), written element by element
    instead of rank_bm25's numpy vectors.  PARITY STATUS: unpinned against rank_bm25 itself (absent);
    pinned by a hand-computed known answer in tests/test_oracle_golden.py.
  * the merge ``dense + bm25 / len(corpus)`` and the stable descending sort (rag_backend.py:790-798).
Only tests/ may import this module.
"""
from __future__ import annotations

import math
from typing import Dict, List, Sequence, Tuple


def bm25_okapi_scores(docs: Sequence[Sequence[str]], query: Sequence[str], k1: float = 1.5, b: float = 0.75,
                      epsilon: float = 0.25) -> List[float]:
    n_docs = len(docs)
    avgdl = sum(len(d) for d in docs) / n_docs
    df: Dict[str, int] = {}
    for d in docs:
        for w in dict.fromkeys(d):  # insertion-ordered set: the idf mean below sums in this order
            df[w] = df.get(w, 0) + 1
    idf = {w: math.log(n_docs - f + 0.5) - math.log(f + 0.5) for w, f in df.items()}
    mean_idf = 0.0
    for w in idf:
        mean_idf += idf[w]
    mean_idf /= len(idf)
    idf = {w: (epsilon * mean_idf if v < 0 else v) for w, v in idf.items()}
    out = []
    for d in docs:
        s = 0.0
        for q in query:
            f = sum(1 for w in d if w == q)
            s += (idf.get(q) or 0) * (f * (k1 + 1) / (f + k1 * (1 - b + b * len(d) / avgdl)))
        out.append(s)
    return out


def hybrid_merge(combined_dense: Dict[str, float], child_docs: Dict[str, str], queries: Sequence[str],
                 max_children: int) -> List[Tuple[str, float]]:
    """rag_backend.py:773-798: BM25 per query variant (max over variants) + dense, stable sort, cut."""
    ids = list(child_docs.keys())
    docs = [child_docs[c].split() for c in ids]
    bm: Dict[str, float] = {}
    if docs:
        for q in queries:
            for c, s in zip(ids, bm25_okapi_scores(docs, q.split())):
                bm[c] = max(bm.get(c, 0.0), float(s))
    merged = {c: d + bm.get(c, 0.0) / (len(docs) or 1) for c, d in combined_dense.items()}
    return sorted(merged.items(), key=lambda it: it[1], reverse=True)[:max_children]
