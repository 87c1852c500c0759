"""Pin the CPU oracle against every golden vector / known answer the reference ships
(SURVEY.md section 8c).  CPU only."""
import hashlib
import itertools

import numpy as np
import pytest

from oracle import cscan, exact_scan as ox, fusion

BGE = "children_baai_bge_small_en_v1_5"
GTE = "children_thenlper_gte_small"
IDS = [
    "217959081514635264", "217959081514635265", "217959081514635266",
    "217959323823771648", "217959323823771649", "217959323823771650",
    "217959323895074816", "217959323895074817", "217959323895074818",
]
# fp64 dots of the stored fp32 vectors (A,B,C = children 1,2,3), SURVEY.md section 8c
KNOWN = {
    BGE: {"AB": 0.9163228374, "AC": 0.8394144321, "BC": 0.8749738904},
    GTE: {"AB": 0.9615456655, "AC": 0.9264357497, "BC": 0.9554065551},
}


def test_fixture_is_the_reference_blob(golden):
    h = hashlib.sha256()
    for r in golden["raw"]["rows"]:
        h.update(bytes.fromhex(r["vector_f32le_hex"]))
    assert h.hexdigest() == "7f7ad93e8959bcbee9363d75aed434e0809abc938919f6ed6d7100ff84cb5e32"
    for name in (BGE, GTE):
        c = golden["collections"][name]
        assert c["ids"] == IDS
        assert c["vectors"].shape == (9, 384)
        nrm = np.linalg.norm(c["vectors"].astype(np.float64), axis=1)
        assert np.all(np.abs(nrm - 1.0) < 2e-7)
        # three ingests of the same three children: rows i, i+3, i+6 are byte-identical
        for i in range(3):
            assert c["blobs"][i] == c["blobs"][i + 3] == c["blobs"][i + 6]
    np.testing.assert_allclose(
        golden["collections"][BGE]["vectors"][0, :4],
        [0.01149409, -0.01299415, -0.0400442, -0.02428864], rtol=0, atol=5e-9)
    np.testing.assert_allclose(
        golden["collections"][GTE]["vectors"][0, :4],
        [-0.02209522, -0.00731409, 0.01469524, -0.01809425], rtol=0, atol=5e-9)
    assert golden["raw"]["collection_metadata"][BGE] == {"key": "hnsw:space", "value": "cosine"}


@pytest.mark.parametrize("name", [BGE, GTE])
def test_known_dot_products(golden, name):
    v = golden["collections"][name]["vectors"].astype(np.float64)
    a, b, c = v[0], v[1], v[2]
    assert abs(a @ b - KNOWN[name]["AB"]) < 1e-9
    assert abs(a @ c - KNOWN[name]["AC"]) < 1e-9
    assert abs(b @ c - KNOWN[name]["BC"]) < 1e-9


@pytest.mark.parametrize("name", [BGE, GTE])
@pytest.mark.parametrize("impl", ["numpy", "c"])
def test_kat_query_is_stored_vector(golden, name, impl):
    """query = stored vector of child A => A1,A2,A3 (score 1), B1,B2,B3, C1,C2,C3; ids in
    insertion order inside each tie group; k=10 over 9 rows returns 9 (+ one pad)."""
    col = golden["collections"][name]
    q = col["vectors"][0:1]
    if impl == "numpy":
        d, rows = ox.exact_topk(q, col["vectors"], 10, "cosine", "f32")
    else:
        d, rows, _ = cscan.exact_topk_prepared(
            ox.prepare_corpus(col["vectors"], "cosine"), ox.prepare_queries(q, "cosine"), 10, "cosine")
    assert rows[0].tolist() == [0, 3, 6, 1, 4, 7, 2, 5, 8, -1]
    score = 1.0 - d[0, :9].astype(np.float64)
    want = [1.0] * 3 + [KNOWN[name]["AB"]] * 3 + [KNOWN[name]["AC"]] * 3
    np.testing.assert_allclose(score, want, rtol=1e-5)
    assert d[0, 0] == d[0, 1] == d[0, 2] and d[0, 3] == d[0, 4] == d[0, 5]
    hits = ox.search_result_dicts(d[0], rows[0], col["ids"], col["metadatas"])
    assert [h["child_id"] for h in hits] == [IDS[i] for i in (0, 3, 6, 1, 4, 7, 2, 5, 8)]
    assert set(hits[0]) == {"score", "child_id", "payload"}
    assert hits[0]["payload"]["parent_id"] == "217959081481080832"


def test_numpy_and_c_oracles_agree_random():
    rng = np.random.default_rng(7)
    c = rng.standard_normal((5000, 384), dtype=np.float32)
    c[100] = c[7]  # planted exact tie
    q = rng.standard_normal((9, 384), dtype=np.float32)
    q[0] = c[7]
    for space in ("cosine", "ip", "l2"):
        for storage in ("f32", "bf16"):
            d0, r0 = ox.exact_topk(q, c, 10, space, storage, chunk_rows=1777)
            pc, pq = ox.prepare_corpus(c, space, storage), ox.prepare_queries(q, space)
            d1, r1, _ = cscan.exact_topk_prepared(pc, pq, 10, space)
            errs = ox.compare_topk_tie_tolerant(
                r1, r0, ox.distances_for_rows(q, c, r1, space, storage), d0.astype(np.float64),
                rtol=1e-5, space=space)
            assert not errs, errs[:3]
            np.testing.assert_allclose(d1, d0, rtol=1e-5, atol=1e-6)
            if space != "l2":
                assert r0[0, 0] == 7 and r0[0, 1] == 100  # tie -> lower row first


def test_bf16_rounding_is_rne():
    x = np.array([1.0, 1.00390625, 1.01171875, -1.00390625, 3.0e-39, 0.0], np.float32)
    # 1 + 2^-8 is exactly halfway between bf16 neighbours 1.0 and 1.0078125 -> ties to even (1.0)
    got = ox.round_to_bf16(x)
    assert got[0] == 1.0 and got[1] == 1.0 and got[3] == -1.0 and got[5] == 0.0
    assert got[2] == np.float32(1.015625)  # 1 + 3*2^-8 halfway -> even mantissa 1.015625
    import torch
    t = torch.randn(10000, dtype=torch.float32)
    np.testing.assert_array_equal(ox.round_to_bf16(t.numpy()), t.bfloat16().float().numpy())


def test_rrf_golden_traces(rrf_traces):
    """Every fused score the reference logged is a bit-exact sum of two 1/(60+rank) terms
    (dual encoder, rag_backend.py:720-731): pins k=60, rank base 1 and fp64 accumulation."""
    assert len(rrf_traces) == 20
    seen = set()
    for tr in rrf_traces:
        for ch in tr["children"]:
            seen.add(ch["retrieval_score"])
    assert len(seen) == 18
    reachable = {}
    for r1, r2 in itertools.combinations_with_replacement(range(1, 31), 2):
        a = [f"x{i}" for i in range(30)]
        a[r1 - 1] = "T"
        b = [f"y{i}" for i in range(30)]
        b[r2 - 1] = "T"
        reachable[dict(fusion.rrf_fuse([a, b], 60))["T"]] = (r1, r2)
    for v in seen:
        assert v in reachable, v


def test_rrf_tie_order_matches_observed_trace(golden):
    """test_logs/query_trace_20250824_121349_f50cc515.json: fused scores 2/61..2/65 -- both
    collections returned the identical copies in insertion order; ties keep first-seen order."""
    col = golden["collections"][BGE]
    d, rows = ox.exact_topk(col["vectors"][0:1], col["vectors"], 5, "cosine")
    lst = [col["ids"][i] for i in rows[0]]
    fused = fusion.rrf_fuse([lst, lst], 60, 5)
    assert [c for c, _ in fused] == [IDS[0], IDS[3], IDS[6], IDS[1], IDS[4]]
    assert [s for _, s in fused] == [1.0 / 61 + 1.0 / 61, 1.0 / 62 + 1.0 / 62, 1.0 / 63 + 1.0 / 63,
                                     1.0 / 64 + 1.0 / 64, 1.0 / 65 + 1.0 / 65]
    a = fusion.rrf_fuse([["p", "q"], ["q", "p"]], 60)
    assert [c for c, _ in a] == ["p", "q"] and a[0][1] == a[1][1]


def test_avg_and_maxsim_restatements():
    f = fusion.avg_fuse([[("a", 0.9), ("b", 0.5), ("c", 0.1)], [("b", 0.7), ("a", 0.7)]])
    assert dict(f) == {"a": 0.5, "b": 0.25, "c": 0.0}
    m = fusion.maxsim_aggregate([[("x", 0.2), ("x", 0.1), ("y", 0.5)], [("y", 0.25)]], 24)
    assert m == [("y", 0.5 + 0.75), ("x", 0.9)]


def test_edge_cases():
    q = np.ones((2, 384), np.float32)
    d, r = ox.exact_topk(q, np.zeros((0, 384), np.float32), 5)
    assert d.shape == (2, 5) and (r == -1).all()
    c = np.zeros((3, 384), np.float32)  # all-zero rows stay finite under cosine
    d, r = ox.exact_topk(q, c, 2)
    assert np.isfinite(d).all() and r.tolist() == [[0, 1], [0, 1]]
    live = np.array([True, False, True])
    d, r = ox.exact_topk(q, c, 3, live=live)
    assert r.tolist() == [[0, 2, -1], [0, 2, -1]]


def test_thresholded_cpu_scan_equals_exact_scan():
    """bench.py's CPU arm (sgemm per chunk + running-threshold compare) returns exactly what the
    plain oracle scan returns, ties and ragged last chunk included."""
    from oracle import exact_scan as ox

    rng = np.random.default_rng(5)
    corpus = rng.standard_normal((5000, 64)).astype(np.float32)
    corpus[4000] = corpus[17]
    corpus[1234] = corpus[17]
    corpus = ox.prepare_corpus(corpus, "cosine", "f32")
    q = ox.prepare_queries(np.concatenate([rng.standard_normal((6, 64)).astype(np.float32), corpus[17:18]]), "cosine")
    for k in (1, 10, 33):
        d0, r0 = ox.exact_topk(q, corpus, k, "cosine", "f32", prepared=True)
        d1, r1 = ox.exact_topk_thresholded(q, corpus, k, "cosine", chunk_rows=700)
        # sgemm over a different chunk shape may round the last bit differently
        assert (r0 == r1).all() and np.abs(d0 - d1).max() <= 5e-7
        # the same scan with the queries dealt to worker threads in slices (the CPU arm at large batches)
        d4, r4 = ox.exact_topk_thresholded_mt(q, corpus, k, "cosine", threads=3, slice_queries=2, chunk_rows=700)
        assert (r0 == r4).all() and np.abs(d0 - d4).max() <= 5e-7
    d2, r2 = ox.exact_topk_thresholded(q, corpus[:7], 10, "cosine", chunk_rows=3)
    d3, r3 = ox.exact_topk(q, corpus[:7], 10, "cosine", "f32", prepared=True)
    assert (r2 == r3).all() and np.allclose(d2, d3, rtol=0, atol=5e-7)


def test_bm25_restatement_known_answer_and_product_host_logic():
    """rank_bm25 is absent (requirements.txt dependency, not vendored): the oracle restates BM25Okapi and is
    pinned by a hand-computed case; the product's host-side BM25 (financial_rag_b200/hybrid.py, numpy
    vectors like rank_bm25) must agree with the oracle's scalar loops bit for bit."""
    import math

    from financial_rag_b200.hybrid import BM25Okapi
    from oracle import bm25 as obm

    docs = [d.split() for d in ["the cat sat on the mat", "the dog ate the cat food", "revenue grew in fiscal 2023",
                                "cat cat cat"]]
    # N = 4, avgdl = 5.  "revenue": df = 1 -> idf = ln 3.5 - ln 1.5; doc 2 has f = 1, dl = avgdl
    #   -> idf * 1 * 2.5 / (1 + 1.5 * 1) = idf.
    s = obm.bm25_okapi_scores(docs, ["revenue"])
    assert s == [0.0, 0.0, math.log(3.5) - math.log(1.5), 0.0]
    # "cat": df = 3 -> idf < 0 -> epsilon * mean idf.  "the": df = 2 -> idf = ln 2.5 - ln 2.5 = 0 (kept).
    idfs = {"cat": math.log(1.5) - math.log(3.5), "the": 0.0}
    for w in ("sat", "on", "mat", "dog", "ate", "food", "revenue", "grew", "in", "fiscal", "2023"):
        idfs[w] = math.log(3.5) - math.log(1.5)
    eps = 0.25 * (sum(idfs.values()) / len(idfs))
    s = obm.bm25_okapi_scores(docs, ["cat"])
    assert s[3] == eps * (3 * 2.5 / (3 + 1.5 * (1 - 0.75 + 0.75 * 3 / 5)))
    assert s[2] == 0.0 and s[0] == s[1] > 0
    rng = np.random.default_rng(0)
    vocab = [f"t{i}" for i in range(40)]
    for _ in range(20):
        corpus = [list(rng.choice(vocab, size=int(rng.integers(1, 30)))) for _ in range(int(rng.integers(1, 25)))]
        query = list(rng.choice(vocab + ["unseen"], size=5))
        assert BM25Okapi(corpus).get_scores(query).tolist() == obm.bm25_okapi_scores(corpus, query)
