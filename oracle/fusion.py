"""CPU ORACLE (test infrastructure, NOT product code) -- cross-collection fusion.

Restates, in plain Python floats (fp64) exactly as the reference computes them:
  * RRF in the retriever            parent_child/retriever.py:82-107
  * RRF in the hybrid backend       rag_backend.py:720-731 (k from ENSEMBLE_RRF_K, default 60)
  * the dead ``avg`` min-max fusion rag_backend.py:732-754
  * MaxSim aggregation              parent_child/multivector_store.py:150-187
Pinned by tests/golden/rrf_traces.json (18 distinct fused scores written by the reference).
Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this module.
"""
from __future__ import annotations

from typing import Dict, List, Sequence, Tuple


def rrf_fuse(
    ranked_lists: Sequence[Sequence[str]], k_rrf: int = 60, top_k: int | None = None
) -> List[Tuple[str, float]]:
    """``agg[cid] += 1.0/(k_rrf+rank)`` with rank from 1, in list order (retriever.py:94-103),
    then Python's stable ``sorted(..., reverse=True)`` (ties keep first-seen order) and ``[:top_k]``
    (retriever.py:104-107).  Empty ids are skipped (retriever.py:98-99)."""
    agg: Dict[str, float] = {}
    for lst in ranked_lists:
        for i, cid in enumerate(lst):
            cid = str(cid or "")
            if not cid:
                continue
            agg[cid] = agg.get(cid, 0.0) + 1.0 / (k_rrf + (i + 1))
    fused = sorted(agg.items(), key=lambda it: it[1], reverse=True)
    return fused if top_k is None else fused[:top_k]


def avg_fuse(
    ranked_lists: Sequence[Sequence[Tuple[str, float]]], top_k: int | None = None
) -> List[Tuple[str, float]]:
    """Per-list min-max normalised scores, summed then divided by the number of lists
    (rag_backend.py:732-754).  A constant list contributes 0 for every member."""
    agg: Dict[str, float] = {}
    nlists = 0
    for lst in ranked_lists:
        nlists += 1  # the reference divides by len(ranked_lists), empty lists included
        scores = [float(s or 0.0) for _, s in lst]
        if not scores:
            continue
        mn, mx = min(scores), max(scores)
        for (cid, _), s in zip(lst, scores):
            cid = str(cid or "")
            if not cid:
                continue
            norm = (s - mn) / (mx - mn) if mx > mn else 0.0
            agg[cid] = agg.get(cid, 0.0) + norm
    if nlists > 0:
        for cid in list(agg.keys()):
            agg[cid] /= float(nlists)
    fused = sorted(agg.items(), key=lambda it: it[1], reverse=True)
    return fused if top_k is None else fused[:top_k]


def maxsim_aggregate(
    per_token_hits: Sequence[Sequence[Tuple[str, float]]], top_k_children: int = 24
) -> List[Tuple[str, float]]:
    """``per_token_hits[t]`` = [(child_id, distance)] of query token t's nearest token vectors.
    Per token: best ``1 - dist`` per child; summed over tokens; stable sort desc; cut
    (multivector_store.py:155-176)."""
    child_scores: Dict[str, float] = {}
    for hits in per_token_hits:
        local_best: Dict[str, float] = {}
        for cid, dist in hits:
            cid = str(cid or "")
            if not cid:
                continue
            score = 1.0 - float(dist)
            if cid not in local_best or score > local_best[cid]:
                local_best[cid] = score
        for cid, s in local_best.items():
            child_scores[cid] = child_scores.get(cid, 0.0) + s
    return sorted(child_scores.items(), key=lambda it: it[1], reverse=True)[:top_k_children]
