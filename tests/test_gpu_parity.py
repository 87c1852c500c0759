"""GPU parity tests: the CUDA path, called through the C ABI, against the CPU oracle.

Run on the B200 box:  python -m pytest tests -m gpu -x -q
Nothing here reads /root/reference; golden data comes from tests/golden/.
"""
import itertools

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")

from oracle import exact_scan as ox  # noqa: E402
from oracle import fusion as ofusion  # noqa: E402

from helpers import (assert_matches_oracle, keys_to_rows, make_corpus, make_queries,  # noqa: E402
                     scores_from_dist)

BGE = "children_baai_bge_small_en_v1_5"
GTE = "children_thenlper_gte_small"
KEY_BASE = 1000


@pytest.fixture(scope="module")
def frb():
    if not torch.cuda.is_available():
        pytest.fail("GPU tests need a CUDA device (there is no CPU fallback to test)")
    import financial_rag_b200 as f

    return f


def build_index(frb, corpus, space, dtype, key_base=KEY_BASE, chunk=None):
    ix = frb.ShardIndex(dim=corpus.shape[1], space=space, dtype=dtype)
    n = corpus.shape[0]
    keys = np.arange(n, dtype=np.int64) + key_base
    step = chunk or max(n, 1)
    for lo in range(0, n, step):
        ix.upsert(corpus[lo:lo + step], keys[lo:lo + step])
    return ix


def stored_rows(ix):
    n = ix.rows()
    if n == 0:
        return np.zeros((0, ix.dim), np.float32)
    return ix.get_rows(0, n)[0]


# ---------------------------------------------------------------------------------------------
# cfg1: the reference's own fixture through the reference-facing store API
class _Child:
    def __init__(self, child_id, parent_id, content, embedding, context=None):
        self.child_id, self.parent_id, self.content = child_id, parent_id, content
        self.embedding, self.context = embedding, context


@pytest.mark.parametrize("dtype", ["f32", "bf16"])
@pytest.mark.parametrize("name", [BGE, GTE])
def test_golden_fixture_through_child_store(frb, golden, name, dtype, monkeypatch, tmp_path):
    monkeypatch.setenv("B200_CHILD_DTYPE", dtype)
    monkeypatch.setenv("CHROMA_CHILD_PERSIST_DIR", str(tmp_path))
    frb.reset_registry()
    col = golden["collections"][name]
    store = frb.get_child_vector_store(collection=name)
    assert store.collection_name == name and store.persist_dir == str(tmp_path)
    assert store.count() == 0
    assert store.search(col["vectors"][0].tolist(), top_k=6) == []
    children = []
    for cid, vec, md in zip(col["ids"], col["vectors"], col["metadatas"]):
        children.append(_Child(int(cid), int(md["parent_id"]), md["snippet"], vec.tolist(), md.get("context")))
    children.append(_Child(1, 2, "no embedding -> skipped", None))
    # three ingests of three children, like the fixture's WAL (seq 1-3, 7-9, 10-12)
    for lo in (0, 3, 6):
        assert store.upsert_children(children[lo:lo + 3] + children[9:]) is True
    assert store.count() == 9
    # a new store object per request sees the same GPU-resident collection (rag_backend.py:632)
    store2 = frb.get_child_vector_store(collection=name)
    hits = store2.search(col["vectors"][0].tolist(), top_k=10)
    want_order = [col["ids"][i] for i in (0, 3, 6, 1, 4, 7, 2, 5, 8)]
    assert [h["child_id"] for h in hits] == want_order
    assert all(set(h) == {"score", "child_id", "payload"} for h in hits)
    assert hits[0]["payload"] == col["metadatas"][0]
    v = col["vectors"].astype(np.float64)
    want = [1.0] * 3 + [float(v[0] @ v[1])] * 3 + [float(v[0] @ v[2])] * 3
    rtol = 1e-5 if dtype == "f32" else 1e-3
    np.testing.assert_allclose([h["score"] for h in hits], want, rtol=rtol)
    # tie groups are bit-identical scores in insertion order
    assert hits[0]["score"] == hits[1]["score"] == hits[2]["score"]
    assert hits[3]["score"] == hits[4]["score"] == hits[5]["score"]
    # torch tensor of shape (1, 384) as the local embedder returns it (retriever.py:87)
    hits_t = store2.search(torch.tensor(col["vectors"][1])[None, :], top_k=6)
    assert [h["child_id"] for h in hits_t][:3] == [col["ids"][i] for i in (1, 4, 7)]
    # upsert of an existing id overwrites in place and keeps its insertion position
    assert store.upsert_children([children[0]]) is True and store.count() == 9
    assert [h["child_id"] for h in store.search(col["vectors"][0], top_k=3)] == want_order[:3]
    frb.reset_registry()


def test_dual_encoder_rrf_on_fixture(frb, golden, tmp_path, monkeypatch):
    """cfg1 + fusion: both collections searched, RRF(60) on the GPU == retriever.py:82-107."""
    monkeypatch.setenv("CHROMA_CHILD_PERSIST_DIR", str(tmp_path))
    monkeypatch.setenv("B200_CHILD_DTYPE", "f32")
    frb.reset_registry()
    lists, key_lists = [], []
    for name in (BGE, GTE):
        col = golden["collections"][name]
        store = frb.get_child_vector_store(collection=name)
        store.upsert_children([_Child(int(c), 7, "t", v.tolist()) for c, v in zip(col["ids"], col["vectors"])])
        hits = store.search(col["vectors"][1].tolist(), top_k=5)
        lists.append([h["child_id"] for h in hits])
        key_lists.append([int(h["child_id"]) for h in hits])
    want = ofusion.rrf_fuse(lists, 60, 5)
    sc, keys = frb.rrf_fuse_host(np.array(key_lists, dtype=np.int64)[:, None, :], 60, 5)
    assert [str(k) for k in keys[0]] == [c for c, _ in want]
    assert sc[0].tolist() == [s for _, s in want]  # bit-exact fp64
    # the observed trace pattern: both encoders agree -> 2/61, 2/62, ...
    assert sc[0].tolist() == [2.0 / 61, 2.0 / 62, 2.0 / 63, 2.0 / 64, 2.0 / 65]
    frb.reset_registry()


# ---------------------------------------------------------------------------------------------
# synthetic corpora vs the oracle
CASES = [
    # n, B, k, space, dtype
    (1, 1, 1, "cosine", "bf16"),
    (9, 1, 10, "cosine", "f32"),
    (31, 3, 10, "cosine", "bf16"),
    (32, 2, 10, "ip", "bf16"),
    (33, 4, 32, "cosine", "bf16"),
    (1000, 5, 10, "cosine", "bf16"),
    (1000, 7, 33, "l2", "bf16"),
    (4097, 1, 50, "cosine", "bf16"),
    (4097, 6, 100, "cosine", "f32"),
    (5003, 9, 128, "ip", "f32"),
    (20000, 4, 10, "l2", "f32"),
    (20000, 8, 10, "cosine", "bf16"),
    (70001, 3, 64, "cosine", "bf16"),
]


@pytest.mark.parametrize("n,B,k,space,dtype", CASES)
def test_scan_matches_oracle(frb, n, B, k, space, dtype):
    dups = [(5, n - 3)] if n > 40 else []
    corpus = make_corpus(n, 384, seed=n, dup_pairs=dups)
    queries = make_queries(B, corpus, seed=B + k)
    if dups:
        queries[0] = corpus[5]  # exact tie between rows 5 and n-3, far apart (different CTAs)
    ix = build_index(frb, corpus, space, dtype, chunk=777 if n > 2000 else None)
    assert ix.count() == n and ix.rows() == n
    dist, keys = ix.search(queries, k)
    rows = keys_to_rows(keys, KEY_BASE)
    assert_matches_oracle(dist, rows, queries, corpus, k, space, dtype, stored=stored_rows(ix),
                          label=f"n={n} B={B} k={k} {space} {dtype}")
    if dups and k >= 2:
        assert rows[0, 0] == 5 and rows[0, 1] == n - 3, "tie must resolve to the lower row first"
        assert dist[0, 0] == dist[0, 1]
    ix.close()


def test_empty_index_and_arg_errors(frb):
    ix = frb.ShardIndex(dim=384, space="cosine", dtype="bf16")
    d, kk = ix.search(np.ones((2, 384), np.float32), 5)
    assert np.isposinf(d).all() and (kk == -1).all()
    d, kk = ix.search(np.zeros((0, 384), np.float32), 5)
    assert d.shape == (0, 5)
    from financial_rag_b200._lib import FrError

    with pytest.raises(FrError):
        ix.search(np.ones((1, 384), np.float32), 129)  # k > FR_MAX_K
    with pytest.raises(FrError):
        ix.search(np.ones((1, 384), np.float32), 0)
    with pytest.raises(ValueError):
        ix.search(np.ones((1, 383), np.float32), 1)
    with pytest.raises(FrError):
        frb.ShardIndex(dim=383)
    ix.close()
    with pytest.raises(RuntimeError):
        ix.count()


@pytest.mark.parametrize("dtype", ["bf16", "f32"])
def test_ingest_kernel_matches_oracle_preparation(frb, dtype):
    corpus = make_corpus(3000, 384, seed=3) * np.float32(3.7)
    corpus[17] = 0.0  # all-zero row stays finite under cosine (1e-30 in the denominator)
    for space in ("cosine", "ip"):
        ix = build_index(frb, corpus, space, dtype)
        got, keys = ix.get_rows(0, 3000)
        assert (keys == np.arange(3000) + KEY_BASE).all()
        want = ox.prepare_corpus(corpus, space, dtype)
        assert np.isfinite(got).all()
        if space == "ip" or dtype == "f32":
            # ip: stored verbatim (bf16: RNE rounding must match bit for bit)
            tol = 0.0 if space == "ip" else 6e-7
            np.testing.assert_allclose(got, want, rtol=tol, atol=0.0 if space == "ip" else 1e-9)
        else:
            # cosine + bf16: the fp32 norm may differ in the last ulp from numpy's summation order,
            # which can flip an RNE decision: at most one bf16 ulp, on a tiny fraction of elements
            diff = np.abs(got - want)
            assert (diff <= np.abs(want) * 2.0 ** -7 + 1e-30).all()
            assert (diff > 0).mean() < 2e-3
        ix.close()


def test_upsert_delete_semantics(frb):
    corpus = make_corpus(600, 384, seed=11)
    ix = frb.ShardIndex(dim=384, space="cosine", dtype="f32")
    keys = np.arange(600, dtype=np.int64) * 3 + 5
    ix.upsert(corpus[:500], keys[:500])
    # duplicate key inside one call: the last vector wins; existing key: overwritten in place
    dup_vecs = np.stack([corpus[510], corpus[511], corpus[512]])
    ix.upsert(dup_vecs, np.array([keys[7], keys[7], keys[501]], dtype=np.int64))
    assert ix.count() == 501 and ix.rows() == 501
    got, gk = ix.get_rows(0, 501)
    model = corpus[:501].copy()
    model[7] = corpus[511]
    model[500] = corpus[512]
    np.testing.assert_allclose(got, ox.prepare_corpus(model, "cosine"), rtol=6e-7, atol=1e-9)
    assert gk[7] == keys[7] and gk[500] == keys[501]
    q = make_queries(4, model, seed=5)
    q[0] = model[7]
    d, kk = ix.search(q, 10)
    assert kk[0, 0] == keys[7]
    # delete: rows disappear from results, count drops, unknown keys are ignored
    victims = np.array([keys[7], keys[100], 999999999], dtype=np.int64)
    assert ix.delete(victims) == 2
    assert ix.count() == 499 and ix.rows() == 501
    live = np.ones(501, bool)
    live[[7, 100]] = False
    d, kk = ix.search(q, 10)
    key_to_row = {int(k): i for i, k in enumerate(gk)}
    rows = np.vectorize(lambda x: key_to_row.get(int(x), -1))(kk)
    assert_matches_oracle(d, rows, q, model, 10, "cosine", "f32", stored=got, live=live, label="after delete")
    assert keys[7] not in kk and keys[100] not in kk
    # re-adding a deleted key appends a new row
    ix.upsert(corpus[7:8], keys[7:8])
    assert ix.count() == 500 and ix.rows() == 502
    d, kk = ix.search(corpus[7:8], 1)
    assert kk[0, 0] == keys[7]
    # deleting everything leaves an empty result
    _, allk = ix.get_rows(0, 502)
    ix.delete(allk[allk != np.iinfo(np.int64).min])
    assert ix.count() == 0
    d, kk = ix.search(q, 3)
    assert (kk == -1).all() and np.isposinf(d).all()
    ix.close()


@pytest.mark.parametrize("dim,dtype", [(768, "bf16"), (768, "f32"), (128, "bf16"), (1024, "f32"), (8, "bf16")])
def test_other_dims(frb, dim, dtype):
    """bf16 x 768 (bert-base tokens, multivector_store.py:70) rides the fast path; the rest the
    generic kernel."""
    corpus = make_corpus(3001, dim, seed=dim)
    q = make_queries(5, corpus, seed=2)
    for space in ("cosine", "l2"):
        ix = build_index(frb, corpus, space, dtype)
        d, kk = ix.search(q, 10)
        assert_matches_oracle(d, keys_to_rows(kk, KEY_BASE), q, corpus, 10, space, dtype, stored=stored_rows(ix),
                              label=f"dim={dim} {dtype} {space}")
        ix.close()


def test_device_api_equals_host_api(frb):
    corpus = make_corpus(50000, 384, seed=21)
    q = make_queries(6, corpus, seed=22)
    ix = build_index(frb, corpus, "cosine", "bf16")
    d, kk = ix.search(q, 10)
    qd = torch.from_numpy(q).cuda()
    dd, dk = ix.search_device(qd, 10)
    torch.cuda.synchronize()
    np.testing.assert_array_equal(dd.cpu().numpy(), d)
    np.testing.assert_array_equal(dk.cpu().numpy(), kk)
    # bulk device ingest == host ingest
    ix2 = frb.ShardIndex(dim=384, space="cosine", dtype="bf16")
    ix2.append_device(torch.from_numpy(corpus).cuda(), None, first_key=KEY_BASE)
    d2, k2 = ix2.search(q, 10)
    np.testing.assert_array_equal(d2, d)
    np.testing.assert_array_equal(k2, kk)
    # upsert after a bulk load rebuilds the key map from the device
    ix2.upsert(corpus[3:4] * 2.0, np.array([KEY_BASE + 3], dtype=np.int64))
    assert ix2.count() == 50000
    ix.close()
    ix2.close()


@pytest.mark.parametrize("G", [2, 3, 8])
def test_row_sharded_merge_equals_single_index(frb, G):
    """K4's merge: G shards searched separately then merged == one index, bit for bit.
    (Shards live on one GPU here; the NCCL all-gather that moves the lists is exercised by
    tests/test_sharded_gloo.py on CPU and by bench.py --gpus N.)"""
    n, B, k = 30011, 5, 10
    corpus = make_corpus(n, 384, seed=31, dup_pairs=[(10, 20000), (11, 29000)])
    q = make_queries(B, corpus, seed=32)
    q[0], q[1] = corpus[10], corpus[11]
    single = build_index(frb, corpus, "cosine", "bf16", key_base=0)
    d1, k1 = single.search(q, k)
    bounds = np.linspace(0, n, G + 1).astype(int)
    packed = torch.zeros((G, B * k), dtype=torch.int64, device="cuda")
    keys = torch.zeros((G, B * k), dtype=torch.int64, device="cuda")
    qd = torch.from_numpy(q).cuda()
    shards = []
    for g in range(G):
        ix = frb.ShardIndex(dim=384, space="cosine", dtype="bf16")
        ix.upsert(corpus[bounds[g]:bounds[g + 1]], np.arange(bounds[g], bounds[g + 1], dtype=np.int64))
        ix.search_partial_device(qd, k, packed[g], keys[g])
        shards.append(ix)
    od = torch.empty((B, k), dtype=torch.float32, device="cuda")
    ok = torch.empty((B, k), dtype=torch.int64, device="cuda")
    frb.merge_shards_device(0, "cosine", packed, keys, B * k, G, B, k, od, ok)
    torch.cuda.synchronize()
    np.testing.assert_array_equal(ok.cpu().numpy(), k1)
    np.testing.assert_array_equal(od.cpu().numpy(), d1)
    assert k1[0, 0] == 10 and k1[0, 1] == 20000  # cross-shard tie -> lower global row
    for ix in shards + [single]:
        ix.close()


def test_rrf_kernel_bit_exact(frb, rrf_traces):
    rng = np.random.default_rng(5)
    for L, B, kp, kout in [(2, 3, 6, 6), (6, 4, 30, 30), (2, 64, 50, 10), (3, 2, 7, 20)]:
        keys = rng.integers(100, 100 + 2 * kp, size=(L, B, kp)).astype(np.int64)
        for l, b in itertools.product(range(L), range(B)):  # unique inside a list, like a k-NN result
            keys[l, b] = rng.permutation(np.arange(100, 100 + 2 * kp))[:kp]
        keys[0, 0, kp - 1] = -1  # a short list
        sc, ok = frb.rrf_fuse_host(keys, 60, kout)
        for b in range(B):
            want = ofusion.rrf_fuse([[str(x) if x >= 0 else "" for x in keys[l, b]] for l in range(L)], 60, kout)
            got = [(str(k_), s) for k_, s in zip(ok[b].tolist(), sc[b].tolist()) if k_ != -1]
            assert got == want
            assert (ok[b, len(want):] == -1).all() and (sc[b, len(want):] == 0).all()
    # device entry point
    kd = torch.from_numpy(keys).cuda()
    sc2, ok2 = frb.rrf_fuse_device(kd, 60, kout)
    torch.cuda.synchronize()
    np.testing.assert_array_equal(sc2.cpu().numpy(), sc)
    np.testing.assert_array_equal(ok2.cpu().numpy(), ok)
    # every fused score the reference logged is reproduced bit-exactly by the kernel
    want_scores = sorted({c["retrieval_score"] for t in rrf_traces for c in t["children"]})
    found = set()
    for r1, r2 in itertools.combinations_with_replacement(range(1, 11), 2):
        a = np.arange(1000, 1010, dtype=np.int64)
        b_ = np.arange(2000, 2010, dtype=np.int64)
        a[r1 - 1], b_[r2 - 1] = 7, 7
        sc, ok = frb.rrf_fuse_host(np.stack([a, b_])[:, None, :], 60, 1)
        assert ok[0, 0] == 7
        found.add(float(sc[0, 0]))
    assert set(want_scores) <= found


def test_multivector_caller_shape(frb, tmp_path):
    """multivector_store.py:142-187 drives Collection.query once per query token and aggregates
    MaxSim on the host; the same aggregation over ONE batched query must agree with the oracle."""
    frb.reset_registry()
    rng = np.random.default_rng(9)
    client = frb.PersistentClient(path=str(tmp_path))
    col = client.get_or_create_collection("parent_child_child_tokens", metadata={"hnsw:space": "cosine"})
    n_child, tok = 40, 12
    vecs = rng.standard_normal((n_child * tok, 768)).astype(np.float32)
    ids = [f"{c}:{t}" for c in range(n_child) for t in range(tok)]
    metas = [{"child_id": str(c), "parent_id": "1", "token_idx": t, "snippet": "s"} for c in range(n_child) for t in range(tok)]
    col.upsert(ids=ids, embeddings=vecs.tolist(), metadatas=metas)
    assert col.count() == n_child * tok
    qtok = vecs[[5, 100, 300, 17]] + 0.05 * rng.standard_normal((4, 768)).astype(np.float32)
    res = col.query(query_embeddings=qtok.tolist(), n_results=10, include=["metadatas", "distances"])
    per_token = [[(m["child_id"], d) for m, d in zip(ms, ds)] for ms, ds in zip(res["metadatas"], res["distances"])]
    got = ofusion.maxsim_aggregate(per_token, 24)
    # oracle: exact scan in fp32, same aggregation
    d, r = ox.exact_topk(qtok, vecs, 10, "cosine", "bf16")
    want = ofusion.maxsim_aggregate([[(str(int(x) // tok), dd) for x, dd in zip(rr, dr)] for rr, dr in zip(r, d)], 24)
    assert [c for c, _ in got][:4] == [c for c, _ in want][:4]
    np.testing.assert_allclose([s for _, s in got], [s for _, s in want], rtol=2e-3)
    # one-by-one calls (what the reference does) give the same lists as the batched call
    one = col.query(query_embeddings=[qtok[2].tolist()], n_results=10, include=["metadatas", "distances", "ids"])
    assert one["ids"][0] == res["ids"][2] and one["distances"][0] == res["distances"][2]
    col.delete(ids=ids[:tok])
    assert col.count() == (n_child - 1) * tok
    frb.reset_registry()


def test_million_rows_gate_with_storage_bound(frb):
    """1M x 384 bf16 Gaussian rows with planted neighbours.  Top scores of the random queries are ~0.25, where a
    RELATIVE 1e-3 is finer than bf16 storage rounding itself, so the gate carries the rigorous storage bound
    (strict=False); the literal gate is applied below on a clustered corpus and measured by bench.py (``parity``)."""
    n, B, k = 1_000_000, 8, 10
    g = torch.Generator(device="cuda").manual_seed(1234)
    c = torch.randn((n, 384), generator=g, device="cuda", dtype=torch.float32)
    ix = frb.ShardIndex(dim=384, space="cosine", dtype="bf16", reserve_rows=n)
    ix.append_device(c, None, first_key=0)
    corpus = c.cpu().numpy()
    q = make_queries(B, corpus, seed=4321)
    d, kk = ix.search(q, k)
    assert_matches_oracle(d, kk, q, corpus, k, "cosine", "bf16", stored=stored_rows(ix), label="1M")
    # planted queries (even indices): their source row is top-1 with a score near 1/sqrt(1+0.01*384)...
    s = scores_from_dist(d, "cosine")
    assert (s[0::2, 0] > 0.4).all() and (s[1::2, 0] < 0.4).all()
    ix.close()


@pytest.mark.parametrize("B,path", [(1, "auto"), (8, "mma"), (64, "mma"), (300, "mma"), (5, "stream")])
def test_clustered_corpus_literal_north_star_gate(frb, B, path):
    """The north-star gate taken literally (strict=True: ids equal except where the fp32 scores tie within 1e-3
    relative, scores within 1e-3 relative, NO storage-bound slack) on the kind of data the reference holds: dense
    neighbourhoods with top scores around 0.9, bf16 storage, every kernel."""
    n, k = 300_000, 10
    corpus, centres = make_clustered(n, 60, 0.35, seed=4100)
    rng = np.random.default_rng(4101 + B)
    queries = (centres[rng.integers(0, 60, B)] + 0.35 / np.sqrt(384.0) * rng.standard_normal((B, 384), dtype=np.float32)).astype(np.float32)
    ix = build_index(frb, corpus, "cosine", "bf16")
    ix.set_path(path)
    d, kk = ix.search(queries, k)
    assert scores_from_dist(d, "cosine")[:, 0].min() > 0.75
    assert_matches_oracle(d, keys_to_rows(kk, KEY_BASE), queries, corpus, k, "cosine", "bf16", strict=True,
                          stored=stored_rows(ix), label=f"literal gate B={B} {path}")
    ix.close()


# ---------------------------------------------------------------------------------------------
# K2: tensor-core selection + exact rescoring + certified fallback
MMA_CASES = [
    # n, B, k
    (1, 1, 1),
    (100, 3, 10),
    (128, 5, 10),
    (129, 64, 16),
    (5000, 128, 10),
    (5000, 129, 17),
    (70001, 200, 32),
    (300000, 7, 10),
    # above 128 queries: CTA pairs (cta_group::2), 256 queries per corpus pass
    (255, 256, 10),
    (257, 130, 3),
    (100000, 300, 10),   # second pass holds 44 queries: the pair's second CTA has none
    (40000, 513, 5),     # two co-resident groups + a tail launch with one query
    (1, 256, 1),
    (60000, 768, 10),    # full launch of two groups, tail launch of one group on all CTA pairs
    (30000, 1030, 8),    # two full launches + tail
    # k' = 128 (k up to 100): candidate lists live in the CTA's slice of the partials array
    (70001, 200, 50),
    (20000, 300, 100),
    (500, 4, 100),
    (90, 2, 100),        # fewer rows than k
    (150000, 20, 64),
]


@pytest.mark.parametrize("n,B,k", MMA_CASES)
def test_mma_path_matches_oracle_and_stream(frb, n, B, k):
    dups = [(5, n - 3)] if n > 40 else []
    corpus = make_corpus(n, 384, seed=1000 + n, dup_pairs=dups)
    queries = make_queries(B, corpus, seed=B + k)
    if dups:
        queries[0] = corpus[5]
    ix = build_index(frb, corpus, "cosine", "bf16")
    ix.set_path("stream")
    d_s, k_s = ix.search(queries, k)
    ix.set_path("mma")
    d_m, k_m = ix.search(queries, k)
    rows = keys_to_rows(k_m, KEY_BASE)
    assert_matches_oracle(d_m, rows, queries, corpus, k, "cosine", "bf16", stored=stored_rows(ix),
                          label=f"mma n={n} B={B} k={k}")
    # the two kernels implement the same definition (fp32 queries on the stored bf16 rows): same ids
    # except fp32-summation-order ties, same scores to accumulation noise
    np.testing.assert_allclose(d_m, d_s, rtol=0, atol=2e-6)
    mism = k_m != k_s
    if mism.any():
        assert np.abs(d_m[mism] - d_s[mism]).max() <= 2e-6
    if dups and k >= 2:
        assert rows[0, 0] == 5 and rows[0, 1] == n - 3
    ix.close()


@pytest.mark.parametrize("n,B,k,dtype", [(300, 3, 10, "bf16"), (70001, 1, 10, "bf16"), (70001, 40, 32, "bf16"),
                                         (50000, 200, 10, "bf16"), (40000, 600, 50, "bf16"), (30000, 130, 100, "bf16"),
                                         (60000, 24, 10, "f32"), (60000, 300, 10, "f32")])
def test_inner_product_collections_on_the_tensor_cores(frb, n, B, k, dtype):
    """The metric vocabulary of the reference (pgvector_child_store.py:7-26: cosine | l2 | ip): inner-product
    collections take the same tensor-core selection + exact rescoring as cosine ones.  Nothing is normalised, so the
    certification bounds scale with |q| * (largest row norm) -- rows of very different lengths, long queries."""
    rng = np.random.default_rng(n + B)
    corpus = make_corpus(n, 384, seed=7000 + n) * rng.uniform(0.05, 3.0, size=(n, 1)).astype(np.float32)
    queries = make_queries(B, corpus, seed=B + k) * rng.uniform(0.2, 5.0, size=(B, 1)).astype(np.float32)
    ix = build_index(frb, corpus, "ip", dtype)
    ix.set_path("stream")
    d_s, k_s = ix.search(queries, k)
    ix.set_path("mma")
    d_m, k_m = ix.search(queries, k)
    assert ix.stat("mma_queries") == B
    rows = keys_to_rows(k_m, KEY_BASE)
    assert_matches_oracle(d_m, rows, queries, corpus, k, "ip", dtype, stored=stored_rows(ix), label=f"ip n={n} B={B} k={k}")
    # same definition in both kernels: scores agree to fp32 summation noise (relative to their size), ids up to such ties
    tol = 3e-6 * np.maximum(np.abs(d_s), 1.0)
    assert (np.abs(d_m - d_s) <= tol)[np.isfinite(d_s)].all()
    mism = (k_m != k_s) & np.isfinite(d_s)
    if mism.any():
        assert (np.abs(d_m - d_s) <= tol)[mism].all()
    ix.set_path("auto")
    d_a, k_a = ix.search(queries, k)
    np.testing.assert_array_equal(k_a, k_m) if n > 2_000_000 or B > 4 else None
    # an overwrite with a much longer row must widen the bounds (the maximum norm is recomputed), never break exactness
    ix.upsert(corpus[:1] * 40.0, np.array([KEY_BASE + 7]))
    corpus2 = corpus.copy()
    corpus2[7] = corpus[0] * 40.0
    ix.set_path("mma")
    d_2, k_2 = ix.search(queries, k)
    assert_matches_oracle(d_2, keys_to_rows(k_2, KEY_BASE), queries, corpus2, k, "ip", dtype, stored=stored_rows(ix),
                          label=f"ip after overwrite n={n} B={B} k={k}")
    ix.close()


@pytest.mark.parametrize("n,B,k", [(300, 3, 10), (70001, 1, 10), (70001, 33, 32), (50000, 64, 10), (40000, 200, 50),
                                   (30000, 40, 100), (30000, 20, 100)])
def test_l2_collections_on_the_tensor_cores(frb, n, B, k):
    """l2 (squared Euclidean, chromadb's definition) through the swapped-operand tensor-core kernel: it ranks rows by
    2 q.c - |c|^2 (one FFMA per score on stored row norms), the exact distances come from direct differences in the rescore
    pass, batches above 64 go through in slices.  Same answer as the CUDA-core stream kernel; clustered rows of mixed
    lengths so that norms matter."""
    rng = np.random.default_rng(n + B)
    corpus, centres = make_clustered(n, 30, 0.5, seed=8000 + n)
    corpus = corpus * rng.uniform(0.5, 2.0, size=(n, 1)).astype(np.float32)
    queries = (centres[rng.integers(0, 30, B)] * rng.uniform(0.5, 2.0, size=(B, 1)) +
               0.3 / np.sqrt(384.0) * rng.standard_normal((B, 384))).astype(np.float32)
    queries[0] = corpus[5]
    ix = build_index(frb, corpus, "l2", "bf16")
    ix.set_path("stream")
    d_s, k_s = ix.search(queries, k)
    ix.set_path("mma")
    d_m, k_m = ix.search(queries, k)
    assert ix.stat("mma_queries") == B
    rows = keys_to_rows(k_m, KEY_BASE)
    assert_matches_oracle(d_m, rows, queries, corpus, k, "l2", "bf16", stored=stored_rows(ix), label=f"l2 n={n} B={B} k={k}")
    fin = np.isfinite(d_s)
    tol = 3e-6 * np.maximum(np.abs(d_s), 1.0)
    assert (np.abs(d_m - d_s) <= tol)[fin].all()
    mism = (k_m != k_s) & fin
    if mism.any():
        assert (np.abs(d_m - d_s) <= tol)[mism].all()
    assert rows[0, 0] == 5 and d_m[0, 0] <= 1e-4
    # overwrite + append keep the stored norms in step with the rows
    ix.upsert(corpus[:2] * 3.0, np.array([KEY_BASE + 9, KEY_BASE + n + 1]))
    corpus2 = np.concatenate([corpus, corpus[1:2] * 3.0, corpus[1:2] * 3.0])[: n + 2]
    corpus2[9] = corpus[0] * 3.0
    d_2, k_2 = ix.search(queries, k)
    rows2 = np.where(k_2 == KEY_BASE + n + 1, n, keys_to_rows(k_2, KEY_BASE))
    assert_matches_oracle(d_2, rows2, queries, corpus2[: n + 1], k, "l2", "bf16", stored=stored_rows(ix),
                          label=f"l2 after mutations n={n} B={B} k={k}")
    ix.close()


@pytest.mark.gpu
def test_mma_co_resident_groups_do_not_change_results(frb):
    """Query groups that share corpus tiles through L2 (option mma_co_groups) are a scheduling
    choice only: ids and distances are bit-identical for 1, 2 and 3 co-resident groups, and the
    uncertified-query counter stays at zero on tie-free data."""
    n, B, k = 50000, 700, 10
    corpus = make_corpus(n, 384, seed=4242)
    queries = make_queries(B, corpus, seed=4243)
    ix = build_index(frb, corpus, "cosine", "bf16")
    ix.set_path("mma")
    ref = None
    for co in (1, 2, 3):
        ix.set_option("mma_co_groups", co)
        d, kk = ix.search(queries, k)
        if ref is None:
            ref = (d, kk)
            assert_matches_oracle(d, keys_to_rows(kk, KEY_BASE), queries, corpus, k, "cosine", "bf16",
                                  stored=stored_rows(ix), label="mma co=1")
        else:
            np.testing.assert_array_equal(kk, ref[1])
            np.testing.assert_array_equal(d, ref[0])
    assert ix.stat("mma_queries") == 3 * B and ix.stat("queries") == 3 * B and ix.stat("searches") == 3
    assert ix.stat("mma_uncertified_queries") == 0
    ix.close()


@pytest.mark.parametrize("B,k", [(2, 10), (16, 10), (17, 16), (32, 32), (33, 10), (64, 16), (64, 32),
                                 # k' = 128 / 256 (cfg3's top-50, cfg5's top-100) and 64 queries at k' = 64: one list per
                                 # query shared by four warps; k' = 256 > 148 CTAs: a CTA publishes its two best rows
                                 (1, 50), (16, 50), (24, 64), (64, 50), (8, 100), (32, 100), (64, 100)])
def test_small_batch_swapped_operand_kernel_equals_k2(frb, B, k):
    """K2s (corpus rows as the MMA's M operand, queries as N; batches <= 64) and K2 select the same candidates:
    after the shared exact rescoring the answers are bit-identical, with and without deleted rows."""
    n = 45000
    corpus = make_corpus(n, 384, seed=900 + B, dup_pairs=[(11, 30000)])
    queries = make_queries(B, corpus, seed=901 + k)
    queries[int(B > 1)] = corpus[11]
    ix = build_index(frb, corpus, "cosine", "bf16")
    ix.set_path("mma")
    for round_ in range(2):
        ix.set_option("mma_small_max", 64)
        d_s, k_s = ix.search(queries, k)
        ix.set_option("mma_small_max", 0)
        d_b, k_b = ix.search(queries, k)
        np.testing.assert_array_equal(k_s, k_b)
        np.testing.assert_array_equal(d_s, d_b)
        live = None
        if round_ == 1:
            live = np.ones(n, bool)
            live[victims - KEY_BASE] = False
        assert_matches_oracle(d_s, keys_to_rows(k_s, KEY_BASE), queries, corpus, k, "cosine", "bf16",
                              stored=stored_rows(ix), live=live, label=f"k2s B={B} k={k} round {round_}")
        if round_ == 0:
            assert keys_to_rows(k_s[int(B > 1)], KEY_BASE)[0] == 11 and (k < 2 or keys_to_rows(k_s[int(B > 1)], KEY_BASE)[1] == 30000)
            victims = np.unique(k_s[:, 0])
            ix.delete(victims)
    ix.close()


@pytest.mark.parametrize("n,B,k", [(700, 3, 100), (3000, 40, 50), (3000, 20, 100), (9000, 64, 32)])
def test_small_batch_kernel_with_fewer_ctas_than_threshold_slots(frb, n, B, k):
    """A corpus of a few tiles runs on a handful of CTAs: each owns several of a query's k' threshold slots and backs
    slot c + r P with its r-th best row (up to 32; beyond that slots stay empty and nothing is shared).  Results are
    the oracle's either way."""
    corpus = make_corpus(n, 384, seed=77 + n)
    queries = make_queries(B, corpus, seed=78 + k)
    ix = build_index(frb, corpus, "cosine", "bf16")
    ix.set_path("mma")
    d, kk = ix.search(queries, k)
    assert_matches_oracle(d, keys_to_rows(kk, KEY_BASE), queries, corpus, k, "cosine", "bf16", stored=stored_rows(ix),
                          label=f"few CTAs n={n} B={B} k={k}")
    assert ix.stat("mma_queries") == B
    ix.close()


@pytest.mark.parametrize("B,k", [(1, 10), (14, 10), (32, 32), (33, 10), (100, 5)])
def test_width_768_on_the_swapped_operand_kernel(frb, B, k):
    """bert-base token vectors (768-d, the multi-vector store's default model, multivector_store.py:70) take K2s
    with a run-time width; batches above what one launch holds (32 queries at this width) go through in slices;
    uncertified queries skip the 384-wide second chance and are re-scanned by the stream kernel."""
    n = 20000
    corpus = make_corpus(n, 768, seed=1700 + B, dup_pairs=[(9, 15000)])
    dup_rows = np.arange(300, 300 + 80 * 100, 100)  # 80 exact copies: more than k' (32 or 64) holds -> stream re-scan
    corpus[dup_rows] = corpus[300]
    queries = make_queries(B, corpus, seed=1701 + k)
    queries[0] = corpus[9]
    if B > 2:
        queries[2] = corpus[300]
    ix = build_index(frb, corpus, "cosine", "bf16")
    ix.set_path("stream")
    d_s, k_s = ix.search(queries, k)
    ix.set_path("mma")
    d_m, k_m = ix.search(queries, k)
    assert ix.stat("mma_queries") == B
    assert_matches_oracle(d_m, keys_to_rows(k_m, KEY_BASE), queries, corpus, k, "cosine", "bf16",
                          stored=stored_rows(ix), label=f"768-d B={B} k={k}")
    np.testing.assert_allclose(d_m, d_s, rtol=0, atol=4e-6)
    mism = k_m != k_s
    if mism.any():
        assert np.abs(d_m[mism] - d_s[mism]).max() <= 4e-6
    if k >= 2:
        assert keys_to_rows(k_m[0], KEY_BASE)[0] == 9 and keys_to_rows(k_m[0], KEY_BASE)[1] == 15000
    if B > 2:
        assert ix.stat("mma_rescanned_queries") >= 1
        np.testing.assert_array_equal(keys_to_rows(k_m[2], KEY_BASE), dup_rows[:k])  # ties in insertion order
    ix.set_path("auto")
    d_a, k_a = ix.search(queries, k)
    np.testing.assert_array_equal(k_a, k_m)
    ix.close()


def test_huge_batch_is_sliced(frb):
    """A batch above the 8192-query slice size is served slice by slice with bounded scratch and equals the
    per-slice answers."""
    n, B, k = 3000, 8192 + 300, 5
    corpus = make_corpus(n, 384, seed=123)
    queries = make_queries(B, corpus, seed=124)
    ix = build_index(frb, corpus, "cosine", "bf16")
    d, kk = ix.search(queries, k)
    d0, k0 = ix.search(queries[:8192], k)
    d1, k1 = ix.search(queries[8192:], k)
    np.testing.assert_array_equal(kk, np.concatenate([k0, k1]))
    np.testing.assert_array_equal(d, np.concatenate([d0, d1]))
    rows = keys_to_rows(kk[8190:8200], KEY_BASE)
    assert_matches_oracle(d[8190:8200], rows, queries[8190:8200], corpus, k, "cosine", "bf16", stored=stored_rows(ix),
                          label="sliced batch")
    ix.close()


@pytest.mark.gpu
@pytest.mark.parametrize("B", [100, 130, 600])
def test_k2_first_tile_threshold_edge_cases(frb, B):
    """K2 (k' = 32) derives a threshold from the maxima of column groups of the first tile before it inserts
    anything.  The cases that bound can get wrong: a query whose scores are all exactly zero (one step below
    +0.0 in key order is -0.0, which compares equal), a first tile of identical rows (every score AT the
    bound), all-negative scores, corpora that end inside the first tile of a stream, tombstones in the first
    tile -- and the option that switches the bound off must not change an answer."""
    k = 10
    for n in (256 * 40 + 17, 256 * 9, 700):
        corpus = make_corpus(n, 384, seed=4000 + n)
        corpus[0:300] = corpus[0]                    # the first tile of stream 0 holds one vector 256 (128) times
        corpus[300:400, 200:] = 0.0                  # rows orthogonal to query 3: exact zeros in a live query's tile
        queries = make_queries(B, corpus, seed=B)
        queries[0] = 0.0                             # every score is +0.0
        queries[1] = corpus[0]
        queries[2] = -corpus[0]
        queries[3] = 0.0
        queries[3, 200:] = corpus[5, 200:]
        ix = build_index(frb, corpus, "cosine", "bf16")
        ix.set_path("stream")
        d_s, k_s = ix.search(queries, k)
        ix.set_path("mma")
        d_m, k_m = ix.search(queries, k)
        ix.set_option("mma_debug", 256 | 512)        # no first-tile bound, compactions one by one
        d_p, k_p = ix.search(queries, k)
        ix.set_option("mma_debug", 0)
        np.testing.assert_array_equal(k_m, k_p)
        np.testing.assert_array_equal(d_m, d_p)
        np.testing.assert_allclose(d_m, d_s, rtol=0, atol=2e-6)
        for i in (0, 1):                             # exact ties: insertion order decides, ids must be the stream path's
            np.testing.assert_array_equal(k_m[i], k_s[i])
        mism = k_m != k_s
        if mism.any():
            assert np.abs(d_m[mism] - d_s[mism]).max() <= 2e-6
        np.testing.assert_array_equal(keys_to_rows(k_m[0], KEY_BASE), np.arange(k))
        np.testing.assert_array_equal(keys_to_rows(k_m[1], KEY_BASE), np.arange(k))
        assert_matches_oracle(d_m, keys_to_rows(k_m, KEY_BASE), queries, corpus, k, "cosine", "bf16",
                              stored=stored_rows(ix), label=f"first tile n={n} B={B}")
        # tombstones inside every stream's first tile: the bound is not used, answers follow the live rows
        dead = np.arange(1, 600, 3) if n > 600 else np.arange(1, n, 3)
        ix.delete((dead + KEY_BASE).astype(np.int64))
        live = np.ones(n, dtype=bool)
        live[dead] = False
        d_d, k_d = ix.search(queries, k)
        assert not np.isin(keys_to_rows(k_d, KEY_BASE), dead).any()
        assert_matches_oracle(d_d, keys_to_rows(k_d, KEY_BASE), queries, corpus, k, "cosine", "bf16",
                              stored=stored_rows(ix), live=live, label=f"first tile with deletes n={n} B={B}")
        ix.close()


def test_mma_certification_fallback_on_mass_ties(frb):
    """More exact duplicates than the k' = 32 selection slots: the tensor-core selection cannot be
    certified, so the query must be re-scanned by the stream kernel and still return the LOWEST
    rows first (insertion order), exactly like the stream path."""
    n, k = 20000, 10
    corpus = make_corpus(n, 384, seed=77)
    dup_rows = np.arange(100, 20000, 150)  # 133 copies of row 100
    corpus[dup_rows] = corpus[100]
    queries = make_queries(6, corpus, seed=78)
    queries[1] = corpus[100]
    queries[4] = corpus[100] + 1e-4 * make_corpus(1, 384, seed=79)[0]
    ix = build_index(frb, corpus, "cosine", "bf16")
    ix.set_path("stream")
    d_s, k_s = ix.search(queries, k)
    ix.set_path("mma")
    d_m, k_m = ix.search(queries, k)
    assert ix.stat("mma_uncertified_queries") >= 2  # queries 1 and 4 failed the first pass ...
    assert ix.stat("mma_rescanned_queries") >= 2    # ... overflowed the second (133 exact ties) and were re-scanned
    np.testing.assert_array_equal(keys_to_rows(k_m[1], KEY_BASE), dup_rows[:k])
    np.testing.assert_array_equal(k_m[1], k_s[1])
    np.testing.assert_array_equal(k_m[4], k_s[4])
    np.testing.assert_allclose(d_m, d_s, rtol=0, atol=2e-6)
    assert_matches_oracle(d_m, keys_to_rows(k_m, KEY_BASE), queries, corpus, k, "cosine", "bf16",
                          stored=stored_rows(ix), label="mma mass ties")
    ix.close()


def test_mma_second_chance_pass_resolves_crowded_neighbourhoods(frb):
    """60 near-duplicates within 3e-4 of the query's best score: more than the k' = 32 selection slots hold, so
    the first tensor-core pass cannot be certified; the second-chance pass (fixed threshold, 128 slots) collects
    the whole crowd and certifies it -- no stream re-scan -- and the answer is the exact fp32-query order."""
    n, k = 30000, 10
    corpus = make_corpus(n, 384, seed=501)
    rng = np.random.default_rng(502)
    crowd = np.arange(200, 200 + 60 * 300, 300)
    for i, r in enumerate(crowd):
        noise = rng.standard_normal(384).astype(np.float32)
        corpus[r] = corpus[100] + 0.024 * (i + 1) / 60.0 * np.linalg.norm(corpus[100]) / np.sqrt(384) * noise
    queries = make_queries(8, corpus, seed=503)
    queries[1] = corpus[100]
    queries[6] = corpus[100] * 0.5  # same direction: cosine ignores the scale
    ix = build_index(frb, corpus, "cosine", "bf16")
    ix.set_path("stream")
    d_s, k_s = ix.search(queries, k)
    ix.set_path("mma")
    ix.set_option("mma_split", 0)  # one-term bf16 queries: selection scores are off by up to ~1e-3
    d_m, k_m = ix.search(queries, k)
    assert ix.stat("mma_uncertified_queries") >= 2, "the crowd must defeat the first pass"
    assert ix.stat("mma_rescanned_queries") == 0, "the second-chance pass must certify it without a stream re-scan"
    # two-term queries (hi + lo) rank the crowd almost exactly: certified in the first pass, same answer
    before = ix.stat("mma_uncertified_queries")
    ix.set_option("mma_split", 1)
    d_2, k_2 = ix.search(queries, k)
    assert ix.stat("mma_uncertified_queries") == before, "two-term selection must certify the crowd at once"
    np.testing.assert_array_equal(k_2, k_m)
    np.testing.assert_array_equal(d_2, d_m)
    ix.set_option("mma_split", 0)
    assert_matches_oracle(d_m, keys_to_rows(k_m, KEY_BASE), queries, corpus, k, "cosine", "bf16",
                          stored=stored_rows(ix), label="mma second chance")
    np.testing.assert_allclose(d_m, d_s, rtol=0, atol=2e-6)
    mism = k_m != k_s
    if mism.any():
        assert np.abs(d_m[mism] - d_s[mism]).max() <= 2e-6
    assert set(keys_to_rows(k_m[1], KEY_BASE)) <= set(crowd.tolist()) | {100}
    # k = 100 with k' = 128 (second chance: 256 slots) on the same crowd
    d_s, k_s = (ix.set_path("stream"), ix.search(queries, 100))[1]
    d_m, k_m = (ix.set_path("mma"), ix.search(queries, 100))[1]
    assert_matches_oracle(d_m, keys_to_rows(k_m, KEY_BASE), queries, corpus, 100, "cosine", "bf16",
                          stored=stored_rows(ix), label="mma second chance k=100")
    np.testing.assert_allclose(d_m, d_s, rtol=0, atol=2e-6)
    ix.close()


def test_scheduling_options_do_not_change_results(frb):
    """Lead throttle between co-resident groups, candidate-list width for k in (64, 100] and the threshold refresh
    schedule only decide WHEN and HOW MUCH is kept, never what the answer is: ids and distances are bit-identical
    across every setting (the exact rescoring sees a superset of the true top-k in all of them)."""
    n, B = 60000, 1100
    corpus = make_corpus(n, 384, seed=6100, dup_pairs=[(17, 41000)])
    queries = make_queries(B, corpus, seed=6101)
    queries[3] = corpus[17]
    ix = build_index(frb, corpus, "cosine", "bf16")
    ix.set_path("mma")
    for k in (10, 100):
        ref = None
        for opts in ({}, {"mma_max_lead": 0}, {"mma_max_lead": 1}, {"mma_co_groups": 2, "mma_max_lead": 2},
                     {"mma_wide_lists": 0}, {"mma_debug": 16}):
            defaults = {"mma_max_lead": 6, "mma_co_groups": 4, "mma_wide_lists": 1, "mma_debug": 0}
            defaults.update(opts)
            for name, value in defaults.items():
                ix.set_option(name, value)
            d, kk = ix.search(queries, k)
            if ref is None:
                ref = (d, kk)
                assert_matches_oracle(d[:24], keys_to_rows(kk[:24], KEY_BASE), queries[:24], corpus, k, "cosine", "bf16",
                                      stored=stored_rows(ix), label=f"scheduling options k={k}")
                assert keys_to_rows(kk[3], KEY_BASE)[0] == 17 and keys_to_rows(kk[3], KEY_BASE)[1] == 41000
            else:
                np.testing.assert_array_equal(kk, ref[1], err_msg=f"k={k} {opts}")
                np.testing.assert_array_equal(d, ref[0], err_msg=f"k={k} {opts}")
    assert ix.stat("mma_rescanned_queries") == 0
    ix.close()


@pytest.mark.parametrize("B,k", [(2, 10), (24, 10), (200, 10), (300, 50)])
def test_f32_collection_batched_search_selects_on_a_bf16_copy(frb, B, k):
    """fp32 collections (the storage that keeps scores within 1e-5): batches run on the tensor cores against a lazily
    built bf16 copy of the rows, every result is rescored on the fp32 rows and certified with the copy's rounding in
    the bound.  Same answer as the CUDA-core stream kernel, through appends, in-place overwrites and deletes."""
    from financial_rag_b200._lib import FrError

    n = 30000
    corpus = make_corpus(n, 384, seed=7100 + B, dup_pairs=[(21, 17000)])
    queries = make_queries(B, corpus, seed=7101 + k)
    queries[1] = corpus[21]
    ix = build_index(frb, corpus[:20000], "cosine", "f32")
    ix.set_option("small_rows_b4", 0)  # keep batch 2 on the tensor-core path for this test
    ix.set_path("mma")
    ix.search(queries, k)                       # builds the copy for the first 20000 rows ...
    ix.upsert(corpus[20000:], np.arange(20000, n, dtype=np.int64) + KEY_BASE)  # ... an append extends it
    for round_ in range(3):
        ix.set_path("stream")
        d_s, k_s = ix.search(queries, k)
        ix.set_path("mma")
        d_m, k_m = ix.search(queries, k)
        assert ix.stat("mma_queries") >= B
        np.testing.assert_allclose(d_m, d_s, rtol=0, atol=2e-6)
        mism = k_m != k_s
        if mism.any():
            assert np.abs(d_m[mism] - d_s[mism]).max() <= 2e-6
        live = None
        if round_ == 2:
            live = np.ones(n, bool)
            live[victims - KEY_BASE] = False
        assert_matches_oracle(d_m, keys_to_rows(k_m, KEY_BASE), queries, corpus, k, "cosine", "f32", stored=stored_rows(ix),
                              live=live, strict=True, label=f"f32 on the tensor cores B={B} k={k} round {round_}")
        if round_ == 0:
            assert keys_to_rows(k_m[1], KEY_BASE)[0] == 21 and keys_to_rows(k_m[1], KEY_BASE)[1] == 17000
            # overwrite a row in place with query 0's direction: the copy must follow
            corpus[12345] = queries[0] * 2.0
            ix.upsert(corpus[12345:12346], np.array([12345 + KEY_BASE], dtype=np.int64))
        elif round_ == 1:
            assert keys_to_rows(k_m[0], KEY_BASE)[0] == 12345
            victims = np.unique(k_m[:, 0])
            ix.delete(victims)
    ix.set_option("mma_f32_shadow", 0)
    with pytest.raises(FrError):
        ix.search(queries, k)  # FR_PATH_MMA without the copy: not served
    ix.close()


def make_clustered(n, centres, spread, seed):
    """Rows = centre + spread * noise: hundreds of rows score within 1e-2 of a query's best hits, like the chunks of
    one document family in a real collection (isotropic Gaussian rows never do)."""
    rng = np.random.default_rng(seed)
    c = rng.standard_normal((centres, 384), dtype=np.float32)
    c /= np.linalg.norm(c, axis=1, keepdims=True)
    which = rng.integers(0, centres, n)
    rows = c[which] + spread / np.sqrt(384.0) * rng.standard_normal((n, 384), dtype=np.float32)
    return rows.astype(np.float32), c


@pytest.mark.parametrize("B", [1, 8, 16])
def test_two_term_queries_certify_a_clustered_corpus_in_one_pass(frb, B):
    """One tight cluster (cosine 0.99 to its centre) puts the k-th and k'-th scores of a query 1e-4 apart, less than
    the one-term selection error even with the score-aware bound (|q - bf16(q)| sqrt(1 - t^2) ~ 1.7e-4): most
    queries fail the first certification.  With the queries read as two bf16 terms (auto for small batches) the
    selection error is ~1e-5 and every query is certified at once; both give the stream kernel's answer."""
    n, k = 200000, 10
    corpus, centres = make_clustered(n, 1, 0.1, seed=3100)
    rng = np.random.default_rng(3101 + B)
    queries = centres[rng.integers(0, 1, B)] + 0.1 / np.sqrt(384.0) * rng.standard_normal((B, 384), dtype=np.float32)
    queries = queries.astype(np.float32)
    ix = build_index(frb, corpus, "cosine", "bf16")
    ix.set_path("stream")
    d_s, k_s = ix.search(queries, k)
    ix.set_path("mma")
    ix.set_option("mma_split", 0)
    d_1, k_1 = ix.search(queries, k)
    fails_one_term = ix.stat("mma_uncertified_queries")
    rescans_one_term = ix.stat("mma_rescanned_queries")  # a crowd this dense can overflow the 128 second-chance slots
    ix.set_option("mma_split", -1)  # auto: two terms at these batch sizes
    d_2, k_2 = ix.search(queries, k)
    assert ix.stat("mma_uncertified_queries") == fails_one_term, "two-term selection left queries uncertified"
    assert ix.stat("mma_rescanned_queries") == rescans_one_term
    if B >= 8:
        assert fails_one_term >= 1, "the corpus is meant to defeat the one-term certification"
    np.testing.assert_allclose(d_1, d_2, rtol=0, atol=2e-6)  # re-scanned queries: the stream kernel's summation order
    mism = k_1 != k_2
    if mism.any():  # ... which may order two rows 1e-7 apart the other way round
        assert np.abs(d_1[mism] - d_2[mism]).max() <= 2e-6 and mism.mean() < 0.05
    np.testing.assert_allclose(d_2, d_s, rtol=0, atol=2e-6)
    mism = k_2 != k_s
    if mism.any():
        assert np.abs(d_2[mism] - d_s[mism]).max() <= 2e-6
    assert_matches_oracle(d_2, keys_to_rows(k_2, KEY_BASE), queries, corpus, k, "cosine", "bf16", strict=False,
                          stored=stored_rows(ix), label=f"clustered two-term B={B}")
    ix.close()


def test_second_chance_blocks_cover_every_failed_query(frb):
    """Hundreds of queries of one batch fail the first certification -- more than one second-chance block of 128
    holds (clustered corpus, error bounds inflated fourfold through the diagnostics option; a larger bound only
    certifies less, the answers stay exact).  Every failure gets its tensor-core second chance, block after block;
    (next to) none is left to the stream re-scan, and the answers equal the stream kernel's."""
    n, B, k = 120000, 700, 10
    corpus, centres = make_clustered(n, 40, 0.3, seed=3200)
    rng = np.random.default_rng(3201)
    queries = centres[rng.integers(0, 40, B)] + 0.3 / np.sqrt(384.0) * rng.standard_normal((B, 384), dtype=np.float32)
    queries = queries.astype(np.float32)
    ix = build_index(frb, corpus, "cosine", "bf16")
    ix.set_path("stream")
    d_s, k_s = ix.search(queries, k)
    ix.set_path("mma")
    d_0, k_0 = ix.search(queries, k)
    assert ix.stat("mma_uncertified_queries") < 70, "the score-aware bound certifies (nearly) the whole batch at once"
    before = ix.stat("mma_uncertified_queries")
    ix.set_option("mma_bound_scale_pct", 400)
    ix.set_option("mma_retry_blocks", -1)  # one second-chance block per 128 queries, whatever the earlier searches needed
    d_m, k_m = ix.search(queries, k)
    fails = ix.stat("mma_uncertified_queries") - before
    assert fails > 128, "need more failures than one second-chance block holds"
    # (with the bound inflated, a query or two may overflow even the 128-slot second-chance list)
    assert ix.stat("mma_rescanned_queries") <= 3, "the second-chance blocks must resolve the failures of every block"
    np.testing.assert_allclose(d_m, d_0, rtol=0, atol=2e-6)
    assert (k_m != k_0).mean() < 0.01
    np.testing.assert_allclose(d_m, d_s, rtol=0, atol=2e-6)
    mism = k_m != k_s
    if mism.any():
        assert np.abs(d_m[mism] - d_s[mism]).max() <= 2e-6
    assert_matches_oracle(d_m[:40], keys_to_rows(k_m[:40], KEY_BASE), queries[:40], corpus, k, "cosine", "bf16",
                          strict=False, stored=stored_rows(ix), label="second-chance blocks")
    ix.close()


def test_second_chance_capacity_follows_the_workload(frb):
    """By default a search enqueues as many second-chance blocks as the last finished search needed (one when nothing
    failed), not one per 128 queries: a 4096-query search used to carry ~100 launches that exit at once.  When more
    queries fail than the blocks hold, the overflow is re-scanned by the stream kernel -- exact all the same -- and
    the next search brings enough blocks."""
    n, B, k = 120000, 700, 10
    corpus, centres = make_clustered(n, 40, 0.3, seed=3200)
    rng = np.random.default_rng(3201)
    queries = (centres[rng.integers(0, 40, B)] + 0.3 / np.sqrt(384.0) * rng.standard_normal((B, 384), dtype=np.float32)).astype(np.float32)
    ix = build_index(frb, corpus, "cosine", "bf16")
    ix.set_path("stream")
    d_s, k_s = ix.search(queries, k)
    ix.set_path("mma")
    ix.set_option("use_graphs", 0)
    from financial_rag_b200 import _lib
    l0 = _lib.launch_count()
    d_0, k_0 = ix.search(queries, k)
    quiet_launches = _lib.launch_count() - l0
    assert ix.stat("mma_retry_blocks") == 1 and quiet_launches <= 16, quiet_launches
    ix.set_option("mma_bound_scale_pct", 400)  # hundreds of first-pass failures from here on
    r0, u0 = ix.stat("mma_rescanned_queries"), ix.stat("mma_uncertified_queries")
    d_1, k_1 = ix.search(queries, k)           # one block enqueued: the overflow takes the stream re-scan
    r1, fails = ix.stat("mma_rescanned_queries"), ix.stat("mma_uncertified_queries") - u0
    assert fails > 128 and r1 - r0 >= fails - 128
    d_2, k_2 = ix.search(queries, k)           # the capacity has followed: (next to) nothing is re-scanned
    assert ix.stat("mma_retry_blocks") >= 2
    assert ix.stat("mma_rescanned_queries") - r1 <= 3
    for d_m, k_m in ((d_1, k_1), (d_2, k_2)):
        np.testing.assert_allclose(d_m, d_s, rtol=0, atol=2e-6)
        mism = k_m != k_s
        if mism.any():
            assert np.abs(d_m[mism] - d_s[mism]).max() <= 2e-6
    ix.set_option("mma_bound_scale_pct", 100)
    ix.search(queries, k)
    ix.search(queries, k)
    assert ix.stat("mma_retry_blocks") == 1    # and it shrinks again when the failures stop
    ix.close()


def test_host_search_graph_replay_and_small_collection_routing(frb):
    """Small collections are launch-bound: the host search replays a captured CUDA graph from the third call of a
    shape on, small batches read / write the pinned block directly, and FR_PATH_AUTO sends batch <= 4 to the 3-launch
    stream kernel.  None of it may change an answer, and an upsert / delete in between must invalidate the graph."""
    n, k = 5000, 10
    corpus = make_corpus(n, 384, seed=777)
    ix = build_index(frb, corpus, "cosine", "bf16")
    for B in (1, 3, 8, 70):
        queries = make_queries(B, corpus, seed=778 + B)
        ix.set_option("use_graphs", 0)
        d_ref, k_ref = ix.search(queries, k)
        ix.set_option("use_graphs", 1)
        r0 = ix.stat("graph_replays")
        for _ in range(4):
            d, kk = ix.search(queries, k)
            np.testing.assert_array_equal(kk, k_ref)
            np.testing.assert_array_equal(d, d_ref)
        assert ix.stat("graph_replays") - r0 >= 2, "calls three and four of a shape must be graph replays"
        # other queries through the same graph
        q2 = make_queries(B, corpus, seed=900 + B)
        d2, k2 = ix.search(q2, k)
        assert_matches_oracle(d2, keys_to_rows(k2, KEY_BASE), q2, corpus, k, "cosine", "bf16", stored=stored_rows(ix),
                              label=f"graph replay B={B}")
    # mutation between replays: the row count / tombstones are baked into the graph, so it must be dropped
    queries = make_queries(3, corpus, seed=781)
    for _ in range(3):
        d, kk = ix.search(queries, k)
    victim = kk[0, 0]
    assert ix.delete(np.array([victim], dtype=np.int64)) == 1
    d, kk = ix.search(queries, k)
    assert victim not in kk[0]
    extra = queries[1:2] * 3.0  # same direction as query 1: becomes its best hit
    ix.upsert(extra, np.array([KEY_BASE + n + 5], dtype=np.int64))
    for _ in range(4):
        d, kk = ix.search(queries, k)
        assert kk[1, 0] == KEY_BASE + n + 5
    # stats count replays like eager searches
    s0, q0 = ix.stat("searches"), ix.stat("queries")
    ix.search(queries, k)
    assert ix.stat("searches") == s0 + 1 and ix.stat("queries") == q0 + 3
    ix.close()


def test_mma_path_with_deletes_and_eligibility(frb):
    from financial_rag_b200._lib import FrError

    n, B, k = 30000, 40, 10
    corpus = make_corpus(n, 384, seed=91)
    queries = make_queries(B, corpus, seed=92)
    ix = build_index(frb, corpus, "cosine", "bf16", key_base=0)
    ix.set_path("mma")
    d0, k0 = ix.search(queries, k)
    victims = np.unique(k0[:, 0])  # delete every current top-1
    assert ix.delete(victims) == len(victims)
    live = np.ones(n, bool)
    live[victims] = False
    d1, k1 = ix.search(queries, k)
    assert not np.isin(k1, victims).any()
    assert_matches_oracle(d1, k1, queries, corpus, k, "cosine", "bf16", stored=stored_rows(ix), live=live,
                          label="mma after delete")
    # auto picks the tensor-core path for this batch and gives the same answer
    ix.set_path("auto")
    d2, k2 = ix.search(queries, k)
    np.testing.assert_array_equal(k2, k1)
    d3, k3 = ix.search(queries, 101)  # k' = 128 leaves no margin: auto serves it with the stream kernel
    np.testing.assert_array_equal(k3[:, :k], k1)
    ix.set_path("mma")
    with pytest.raises(FrError):
        ix.search(queries, 101)
    ix.close()
    for kw in ({"dtype": "f32", "dim": 768}, {"space": "l2", "dtype": "f32"}, {"space": "l2", "dim": 768}, {"dim": 128}):
        args = {"dim": 384, "space": "cosine", "dtype": "bf16"}
        args.update(kw)
        jx = frb.ShardIndex(**args)
        jx.upsert(make_corpus(10, args["dim"], seed=1), np.arange(10))
        jx.set_path("mma")
        with pytest.raises(FrError):
            jx.search(make_corpus(1, args["dim"], seed=2), 5)
        jx.set_path("auto")  # auto silently uses the kernel that serves the configuration
        jx.search(make_corpus(8, args["dim"], seed=2), 5)
        jx.close()


def test_concurrent_host_searches_share_a_shard(frb):
    """Flask serves requests from several threads (api_server.py:1366-1371): concurrent fr_index_search calls on one
    shard hold its lock only while they are enqueued, wait for their own event and read their own pinned slot.
    Every thread must get exactly the answer of a lone call -- small batches (graph replay, zero-copy), large ones
    (copies through the staging buffers), more threads than slots."""
    import threading

    n = 60000
    corpus = make_corpus(n, 384, seed=9100)
    ix = build_index(frb, corpus, "cosine", "bf16")
    jobs = []
    for t in range(7):
        b = (1, 3, 8, 40, 70, 200, 300)[t]
        q = make_queries(b, corpus, seed=9200 + t)
        jobs.append((q, ix.search(q, 10)))           # the lone answers (also warms the graphs up)
    errors = []

    def worker(t):
        q, (want_d, want_k) = jobs[t]
        try:
            for _ in range(25):
                d, kk = ix.search(q, 10)
                if not ((kk == want_k).all() and (d == want_d).all()):
                    errors.append(f"thread {t}: answer differs from the lone call")
                    return
        except Exception as e:  # noqa: BLE001
            errors.append(f"thread {t}: {e!r}")

    threads = [threading.Thread(target=worker, args=(t,)) for t in range(7)]
    for th in threads:
        th.start()
    for th in threads:
        th.join()
    assert not errors, errors
    assert ix.stat("searches") >= 7 * 26
    ix.close()


@pytest.mark.gpu
def test_concurrent_host_searches_in_a_fresh_process(frb):
    """The same load in a process of its own (scripts/stress_concurrent.py), one shard and two: there the first graph
    captures and replays of the process meet other threads' waits -- the combination that crashed inside cuGraphLaunch
    before graph work and host waits excluded each other (fr_host.h: graph_wait_mutex)."""
    import os
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for extra in ({}, {"FR_STRESS_TWO": "1"}):
        env = dict(os.environ, FRB200_SEGV_TRACE="1", **extra)
        r = subprocess.run([sys.executable, os.path.join(root, "scripts", "stress_concurrent.py"), "8"], env=env,
                           capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, (r.returncode, r.stdout[-500:], r.stderr[-1500:])
