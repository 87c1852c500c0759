"""Shared checkers for the parity tests: CUDA path vs the CPU oracle (oracle/).

Two gates, applied together:

(1) NORTH-STAR GATE against the fp32 exact scan of the ORIGINAL fp32 vectors (the reference's
    definition): ids identical except where the two candidates' oracle scores tie within 1e-3
    relative; scores within 1e-3 relative (bf16 storage) or 1e-5 (fp32 storage).
    With ``strict=True`` exactly that.  With ``strict=False`` (synthetic Gaussian data, where
    top scores are ~0.1 and a *relative* 1e-3 is smaller than bf16 rounding itself) the
    rigorous bf16 storage bound  2^-9 * sum_i |q_i c_i|  is added as absolute slack -- any
    correct bf16 index needs it; it is zero for fp32 storage.
(2) STORAGE-EXACT GATE against the exact scan of the values the index really holds (read back
    through fr_index_get_rows): only fp32 summation-order noise is allowed (1e-5 relative),
    for both storage types.  This is the check that catches kernel bugs.
"""
from __future__ import annotations

import numpy as np

from oracle import exact_scan as ox

SCORE_RTOL = {"bf16": 1e-3, "f32": 1e-5}
ID_TIE_RTOL = 1e-3
# fp32 accumulation of <= 4096 products has an absolute error floor of a few 1e-7 * sum|q_i c_i|;
# a purely relative bound is vacuous for scores near 0 (orthogonal vectors).
ACC_ATOL = 2e-6


def scores_from_dist(dist: np.ndarray, space: str) -> np.ndarray:
    """The number the reference reports: ``1 - dist`` (chroma_child_store.py:72); l2 compares the
    distance itself."""
    d = np.asarray(dist, dtype=np.float64)
    return d if ox.canonical_space(space) == "l2" else 1.0 - d


def keys_to_rows(keys: np.ndarray, key_base: int = 0) -> np.ndarray:
    rows = keys.astype(np.int64) - key_base
    rows[keys == -1] = -1
    return rows


def _pair_stats(q_prep: np.ndarray, c_prep: np.ndarray, rows: np.ndarray, space: str):
    """(distance in fp64, bf16 storage bound) of query i vs prepared corpus rows[i, j]."""
    dist = np.full(rows.shape, np.inf)
    bound = np.zeros(rows.shape)
    q64 = q_prep.astype(np.float64)
    for i in range(rows.shape[0]):
        ok = rows[i] >= 0
        if not ok.any():
            continue
        c = c_prep[rows[i, ok]].astype(np.float64)
        if space == "l2":
            diff = c - q64[i][None, :]
            dist[i, ok] = np.einsum("ij,ij->i", diff, diff)
            # (c(1+e) - q)^2 - (c - q)^2 = 2 c e (c - q) + c^2 e^2 with |e| <= 2^-9
            bound[i, ok] = 2.0 ** -9 * 2.0 * np.abs(c * diff).sum(axis=1) + 2.0 ** -18 * (c * c).sum(axis=1)
        else:
            dist[i, ok] = 1.0 - c @ q64[i]
            bound[i, ok] = 2.0 ** -9 * np.abs(c * q64[i][None, :]).sum(axis=1)
    return dist, bound


def _id_gate(got_rows, ref_rows, s_got, s_ref, rtol, slack_got, slack_ref, label):
    errs = []
    for i in range(ref_rows.shape[0]):
        for j in range(ref_rows.shape[1]):
            g, r = int(got_rows[i, j]), int(ref_rows[i, j])
            if g == r:
                continue
            if g < 0 or r < 0:
                errs.append(f"q{i} pos{j}: row {g} vs oracle {r} (length mismatch)")
                continue
            a, b = s_got[i, j], s_ref[i, j]
            if abs(a - b) > rtol * max(abs(a), abs(b)) + slack_got[i, j] + slack_ref[i, j] + 1e-12:
                errs.append(f"q{i} pos{j}: row {g} (oracle score {a:.7f}) vs oracle row {r} ({b:.7f})")
    assert not errs, f"{label}: id gate failed ({len(errs)}): " + "; ".join(errs[:3])


def assert_matches_oracle(got_dist, got_rows, queries, corpus_f32, k, space, storage, *, stored=None,
                          live=None, strict=False, label=""):
    space = ox.canonical_space(space)
    b = queries.shape[0]
    assert got_dist.shape == (b, k) and got_rows.shape == (b, k), (got_dist.shape, got_rows.shape)
    ok = got_rows >= 0
    assert np.all(np.isposinf(got_dist[~ok])), f"{label}: pad distances must be +inf"
    q_prep = ox.prepare_queries(queries, space)

    # ---- (1) north-star gate: the reference's definition on the original fp32 vectors
    c_prep = ox.prepare_corpus(corpus_f32, space, "f32")
    ref_d, ref_r = ox.exact_topk(q_prep, c_prep, k, space, "f32", prepared=True, live=live)
    assert ((got_rows >= 0).sum(axis=1) == (ref_r >= 0).sum(axis=1)).all(), f"{label}: result lengths differ"
    d_got, bound_got = _pair_stats(q_prep, c_prep, got_rows, space)
    d_ref, bound_ref = _pair_stats(q_prep, c_prep, ref_r, space)
    if strict or storage == "f32":
        bound_got = np.zeros_like(bound_got)
        bound_ref = np.zeros_like(bound_ref)
    s_got, s_ref = scores_from_dist(d_got, space), scores_from_dist(d_ref, space)
    _id_gate(got_rows, ref_r, s_got, s_ref, ID_TIE_RTOL, bound_got, bound_ref, label + " [north-star]")
    a, r = scores_from_dist(got_dist, space)[ok], s_got[ok]
    tol = SCORE_RTOL[storage] * np.maximum(np.abs(a), np.abs(r)) + ACC_ATOL + bound_got[ok]
    bad = np.abs(a - r) > tol
    assert not bad.any(), (f"{label}: {bad.sum()} scores outside the north-star tolerance; first: got {a[bad][0]:.7g} "
                           f"oracle {r[bad][0]:.7g} tol {tol[bad][0]:.3g}")

    # ---- (2) storage-exact gate: exact scan of what the index holds
    if stored is not None:
        sd, sr = ox.exact_topk(q_prep, stored, k, space, "f32", prepared=True, live=live)
        d_got2, _ = _pair_stats(q_prep, stored, got_rows, space)
        d_ref2, _ = _pair_stats(q_prep, stored, sr, space)
        zero = np.zeros_like(d_got2)
        _id_gate(got_rows, sr, scores_from_dist(d_got2, space), scores_from_dist(d_ref2, space), 1e-5,
                 zero + ACC_ATOL, zero, label + " [storage-exact]")
        a, r = scores_from_dist(got_dist, space)[ok], scores_from_dist(d_got2, space)[ok]
        bad = np.abs(a - r) > 1e-5 * np.maximum(np.abs(a), np.abs(r)) + ACC_ATOL
        assert not bad.any(), f"{label}: storage-exact scores off, worst {np.abs(a - r).max():.3e}"


def make_corpus(n, dim=384, seed=0, dup_pairs=()):
    rng = np.random.default_rng(seed)
    c = rng.standard_normal((n, dim), dtype=np.float32)
    for a, b in dup_pairs:
        c[b] = c[a]
    return c


def make_queries(b, corpus, seed=1, planted=True):
    rng = np.random.default_rng(seed)
    q = rng.standard_normal((b, corpus.shape[1]), dtype=np.float32)
    if planted and corpus.shape[0] > 0:
        for i in range(0, b, 2):  # half the queries are noisy copies of corpus rows
            j = int(rng.integers(0, corpus.shape[0]))
            q[i] = corpus[j] + 0.1 * rng.standard_normal(corpus.shape[1]).astype(np.float32)
    return q
