"""CPU ORACLE (test infrastructure, NOT product code) -- exact child-chunk similarity scan.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference``
legs may import this module.  The product path (``financial_rag_b200``) never does; it fails
loudly when the CUDA extension is missing.

PARITY STATUS: **partially pinned**.  The arithmetic of the reference's path lives in the
third-party wheel ``chromadb`` (``requirements.txt:21`` pins only ``chromadb>=0.5.4``; the shipped
fixture was written by a 1.x Rust-core build) which is absent from ``/root/reference`` and from
this image, and the reference has no test that asserts a search result -- so for the ranked result
list of the chromadb call itself this oracle is **parity unpinned** (neither a reference golden vector nor a
runnable reference exists for it).  What IS pinned on the reference's own artefacts:
  * the 18 golden fp32 vectors in Chroma's WAL (``tests/golden/chroma_fixture.json``) and the
    known answers derived from them;
  * the 18 distinct fused scores in ``test_logs/query_trace_*.json`` (``tests/golden/rrf_traces.json``),
    all bit-exact sums of ``1/(60+rank)``;
  * the tie order observed in ``test_logs/query_trace_20250824_121349_f50cc515.json``
    (identical vectors come back in insertion order).
The distance definitions themselves are a restatement of chromadb/hnswlib's published behaviour:

  cosine : rows and queries L2-normalised as x * 1/(||x|| + 1e-30), d = 1 - sum(a_i b_i)   (fp32)
  l2     : d = sum((a_i - b_i)^2)   (squared, no sqrt)                                       (fp32)
  ip     : d = 1 - sum(a_i b_i), no normalisation                                            (fp32)

Result order (restating ``ChromaChildStore.search``, parent_child/chroma_child_store.py:62-74):
ascending distance, ties by ascending insertion row; ``score = 1.0 - float(dist)``.
HNSW is approximate; this oracle is the exact scan the north star names as the reference.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Tuple

import os

import numpy as np

SPACES = ("cosine", "l2", "ip")


def canonical_space(space: Optional[str]) -> str:
    """Metric vocabulary of parent_child/pgvector_child_store.py:7-26 (cosine is the default)."""
    d = (space or "cosine").lower()
    if d in ("cos", "cosine"):
        return "cosine"
    if d in ("l2", "euclidean"):
        return "l2"
    if d in ("ip", "inner", "inner_product"):
        return "ip"
    return "cosine"


def normalize_rows(x: np.ndarray) -> np.ndarray:
    """x * 1/(||x|| + 1e-30), all in fp32 (hnswlib-style cosine normalisation)."""
    x = np.ascontiguousarray(x, dtype=np.float32)
    if x.ndim == 1:
        x = x[None, :]
    # fp32 sum of squares (pairwise in numpy); the 1e-30 keeps all-zero rows finite.
    nrm = np.sqrt(np.einsum("ij,ij->i", x, x, dtype=np.float32)).astype(np.float32)
    inv = (np.float32(1.0) / (nrm + np.float32(1e-30))).astype(np.float32)
    return (x * inv[:, None]).astype(np.float32)


def round_to_bf16(x: np.ndarray) -> np.ndarray:
    """Round fp32 -> bf16 (round-to-nearest-even) and return the values as fp32."""
    x = np.ascontiguousarray(x, dtype=np.float32)
    u = x.view(np.uint32).astype(np.uint64)
    u = (u + 0x7FFF + ((u >> 16) & 1)) & 0xFFFF0000
    return u.astype(np.uint32).view(np.float32).reshape(x.shape)


def prepare_corpus(vectors: np.ndarray, space: str = "cosine", storage: str = "f32") -> np.ndarray:
    """What the index stores for a batch of inserted vectors (K0 in DESIGN.md)."""
    space = canonical_space(space)
    c = np.ascontiguousarray(vectors, dtype=np.float32)
    if c.ndim == 1:
        c = c[None, :]
    if space == "cosine":
        c = normalize_rows(c)
    if storage == "bf16":
        c = round_to_bf16(c)
    elif storage != "f32":
        raise ValueError(storage)
    return c


def prepare_queries(queries: np.ndarray, space: str = "cosine") -> np.ndarray:
    space = canonical_space(space)
    q = np.ascontiguousarray(queries, dtype=np.float32)
    if q.ndim == 1:
        q = q[None, :]
    if space == "cosine":
        q = normalize_rows(q)
    return q


def distances(q: np.ndarray, c: np.ndarray, space: str = "cosine", acc=np.float32) -> np.ndarray:
    """[B, N] distances between PREPARED queries and PREPARED corpus rows."""
    space = canonical_space(space)
    if space in ("cosine", "ip"):
        dot = (q.astype(acc) @ c.astype(acc).T).astype(acc)
        return (acc(1.0) - dot).astype(np.float32)
    # l2: squared Euclidean, direct differences (no norm expansion, no cancellation)
    out = np.empty((q.shape[0], c.shape[0]), dtype=np.float32)
    for i in range(q.shape[0]):
        diff = c.astype(acc) - q[i].astype(acc)[None, :]
        out[i] = np.einsum("ij,ij->i", diff, diff).astype(np.float32)
    return out


def _select_topk_rows(d: np.ndarray, k: int, row_base: int) -> Tuple[np.ndarray, np.ndarray]:
    """Exact k smallest of each row of d under (dist asc, row asc); returns (dist, rows)."""
    b, n = d.shape
    kk = min(k, n)
    out_d = np.full((b, k), np.inf, dtype=np.float32)
    out_r = np.full((b, k), -1, dtype=np.int64)
    if kk == 0:
        return out_d, out_r
    if kk < n:
        kth = np.partition(d, kk - 1, axis=1)[:, kk - 1]
    else:
        kth = d.max(axis=1)
    for i in range(b):
        cand = np.nonzero(d[i] <= kth[i])[0]  # every tie at the boundary is a candidate
        order = np.lexsort((cand, d[i, cand]))[:kk]
        sel = cand[order]
        out_d[i, :kk] = d[i, sel]
        out_r[i, :kk] = sel + row_base
    return out_d, out_r


def merge_topk(
    parts: Sequence[Tuple[np.ndarray, np.ndarray]], k: int
) -> Tuple[np.ndarray, np.ndarray]:
    """Merge partial (dist, rows) lists -> top-k under (dist asc, row asc); pads with (inf, -1)."""
    d = np.concatenate([p[0] for p in parts], axis=1)
    r = np.concatenate([p[1] for p in parts], axis=1)
    b = d.shape[0]
    out_d = np.full((b, k), np.inf, dtype=np.float32)
    out_r = np.full((b, k), -1, dtype=np.int64)
    for i in range(b):
        valid = np.nonzero(r[i] >= 0)[0]
        order = valid[np.lexsort((r[i, valid], d[i, valid]))][:k]
        out_d[i, : len(order)] = d[i, order]
        out_r[i, : len(order)] = r[i, order]
    return out_d, out_r


def exact_topk(
    queries: np.ndarray,
    corpus: np.ndarray,
    k: int,
    space: str = "cosine",
    storage: str = "f32",
    chunk_rows: int = 1 << 20,
    prepared: bool = False,
    live: Optional[np.ndarray] = None,
) -> Tuple[np.ndarray, np.ndarray]:
    """Exact top-k scan: returns (dist [B,k] fp32, rows [B,k] int64, -1 padded).

    ``corpus`` rows are in insertion order; ``live`` (bool [N]) masks deleted rows.
    With ``prepared=True`` both inputs are taken as already normalised / rounded.
    """
    space = canonical_space(space)
    q = np.ascontiguousarray(queries, dtype=np.float32) if prepared else prepare_queries(queries, space)
    if q.ndim == 1:
        q = q[None, :]
    n = corpus.shape[0]
    parts = []
    for lo in range(0, max(n, 1), chunk_rows):
        blk = corpus[lo : lo + chunk_rows]
        if blk.shape[0] == 0:
            break
        c = np.ascontiguousarray(blk, dtype=np.float32) if prepared else prepare_corpus(blk, space, storage)
        d = distances(q, c, space)
        if live is not None:
            d = np.where(live[lo : lo + chunk_rows][None, :], d, np.float32(np.inf))
        pd, pr = _select_topk_rows(d, k, lo)
        if live is not None:
            pr = np.where(np.isfinite(pd), pr, -1)
        parts.append((pd, pr))
    if not parts:
        b = q.shape[0]
        return np.full((b, k), np.inf, np.float32), np.full((b, k), -1, np.int64)
    return merge_topk(parts, k)


def exact_topk_thresholded(
    queries: np.ndarray, corpus: np.ndarray, k: int, space: str = "cosine", chunk_rows: int = 1 << 16
) -> Tuple[np.ndarray, np.ndarray]:
    """Same result as ``exact_topk(..., prepared=True)``, organised the way a CPU scan is fast at
    large batches: one sgemm per corpus chunk, then a vectorised compare against each query's
    running k-th distance so only the few rows that can still enter a list are touched.
    Used by bench.py's CPU arm; checked against ``exact_topk`` in tests/test_oracle_golden.py."""
    space = canonical_space(space)
    q = np.ascontiguousarray(queries, dtype=np.float32)
    b, n = q.shape[0], corpus.shape[0]
    best_d = np.full((b, k), np.inf, dtype=np.float32)
    best_r = np.full((b, k), -1, dtype=np.int64)
    for lo in range(0, n, chunk_rows):
        c = np.ascontiguousarray(corpus[lo : lo + chunk_rows], dtype=np.float32)
        d = distances(q, c, space)
        if lo == 0 or not np.isfinite(best_d[:, -1]).all():
            pd, pr = _select_topk_rows(d, k, lo)
            best_d, best_r = merge_topk([(best_d, best_r), (pd, pr)], k)
            continue
        qi, ri = np.nonzero(d <= best_d[:, -1][:, None])  # ties at the boundary stay candidates
        if qi.size == 0:
            continue
        starts = np.flatnonzero(np.r_[True, qi[1:] != qi[:-1]])
        ends = np.r_[starts[1:], qi.size]
        for s0, e0 in zip(starts, ends):
            i = int(qi[s0])
            cd = np.concatenate([best_d[i], d[i, ri[s0:e0]]])
            cr = np.concatenate([best_r[i], ri[s0:e0] + lo])
            valid = np.nonzero(cr >= 0)[0]
            order = valid[np.lexsort((cr[valid], cd[valid]))][:k]
            best_d[i, : len(order)] = cd[order]
            best_r[i, : len(order)] = cr[order]
    return best_d, best_r


def exact_topk_thresholded_mt(
    queries: np.ndarray, corpus: np.ndarray, k: int, space: str = "cosine", threads: int = 0,
    slice_queries: int = 128, chunk_rows: int = 1 << 14
) -> Tuple[np.ndarray, np.ndarray]:
    """``exact_topk_thresholded`` with the queries dealt to ``threads`` worker threads in slices (the lists of
    different queries never meet, so the slices are independent scans).  One big sgemm per chunk leaves the
    element-wise passes behind it -- ``1 - dot``, the compare against the running k-th distances, ``nonzero`` -- on one
    core and they take twice the sgemm's time; per slice they run on every core, on blocks that stay in cache
    (numpy releases the GIL in all of them).  The caller sets the BLAS pool to ONE thread per call (threadpoolctl).
    Same results, bit for bit (tests/test_oracle_golden.py); 4 x the throughput on 8 cores at batch 4096."""
    from concurrent.futures import ThreadPoolExecutor

    q = np.ascontiguousarray(queries, dtype=np.float32)
    b = q.shape[0]
    threads = threads or (os.cpu_count() or 1)
    slices = [(i, min(i + slice_queries, b)) for i in range(0, b, slice_queries)]
    if len(slices) <= 1 or threads <= 1:
        return exact_topk_thresholded(q, corpus, k, space, chunk_rows=chunk_rows)

    def work(se):
        return exact_topk_thresholded(q[se[0] : se[1]], corpus, k, space, chunk_rows=chunk_rows)

    with ThreadPoolExecutor(max_workers=threads) as ex:
        parts = list(ex.map(work, slices))
    return np.concatenate([p[0] for p in parts]), np.concatenate([p[1] for p in parts])


def distances_for_rows(
    queries: np.ndarray, corpus: np.ndarray, rows: np.ndarray, space: str = "cosine", storage: str = "f32"
) -> np.ndarray:
    """fp64-accumulated distance of query i to corpus rows[i, j] (for the tie-tolerant comparator)."""
    space = canonical_space(space)
    q = prepare_queries(queries, space).astype(np.float64)
    out = np.full(rows.shape, np.inf, dtype=np.float64)
    for i in range(rows.shape[0]):
        ok = rows[i] >= 0
        if not ok.any():
            continue
        c = prepare_corpus(corpus[rows[i, ok]], space, storage).astype(np.float64)
        if space == "l2":
            diff = c - q[i][None, :]
            out[i, ok] = np.einsum("ij,ij->i", diff, diff)
        else:
            out[i, ok] = 1.0 - c @ q[i]
    return out


def compare_topk_tie_tolerant(
    got_rows: np.ndarray,
    ref_rows: np.ndarray,
    ref_dist_of_got: np.ndarray,
    ref_dist: np.ndarray,
    rtol: float = 1e-3,
    space: str = "cosine",
) -> List[str]:
    """The north star's id gate: ids identical to the fp32 exact scan except where the two
    candidates' REFERENCE scores tie within ``rtol`` relative.  ``ref_dist_of_got[i,j]`` is the
    oracle distance of the row the kernel returned at (i,j).  Returns a list of violations.

    "score" is ``1 - dist`` for cosine/ip (the number the reference returns,
    chroma_child_store.py:72) and the distance itself for l2.
    """
    space = canonical_space(space)
    errs = []
    for i in range(ref_rows.shape[0]):
        for j in range(ref_rows.shape[1]):
            g, r = int(got_rows[i, j]), int(ref_rows[i, j])
            if g == r:
                continue
            if g < 0 or r < 0:
                errs.append(f"query {i} pos {j}: got row {g}, oracle row {r} (length mismatch)")
                continue
            a, b = float(ref_dist_of_got[i, j]), float(ref_dist[i, j])
            if space != "l2":
                a, b = 1.0 - a, 1.0 - b
            if abs(a - b) > rtol * max(abs(a), abs(b)) + 1e-12:
                errs.append(
                    f"query {i} pos {j}: got row {g} (oracle score {a:.7f}) vs oracle row {r} "
                    f"(score {b:.7f}): not a tie within {rtol}"
                )
    return errs


def search_result_dicts(
    dist_row: np.ndarray, rows_row: np.ndarray, ids: Sequence[str], metadatas: Sequence[Optional[dict]]
) -> List[Dict]:
    """Result shape of ChromaChildStore.search (parent_child/chroma_child_store.py:62-74)."""
    out = []
    for d, r in zip(dist_row, rows_row):
        if r < 0:
            continue
        out.append({"score": 1.0 - float(d), "child_id": ids[int(r)], "payload": metadatas[int(r)] or {}})
    return out
