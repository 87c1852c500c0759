"""GPU parity tests for the SURVEY.md 8(f) rows: persistence (f-2), multi-vector MaxSim on the GPU (f-3),
the document-level helpers the reference probes for, and the BM25 + dense merge (f-1)."""
from __future__ import annotations

import json
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

import torch  # noqa: E402

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from oracle import exact_scan as ox  # noqa: E402
from oracle import fusion as ofusion  # noqa: E402

from helpers import make_corpus, make_queries  # noqa: E402


@pytest.fixture(scope="module")
def frb():
    if not torch.cuda.is_available():
        pytest.fail("GPU tests need a CUDA device (there is no CPU fallback to test)")
    import financial_rag_b200 as f

    return f


class _Child:
    def __init__(self, child_id, parent_id, content, embedding=None, context=None, document_id=None):
        self.child_id, self.parent_id, self.content = child_id, parent_id, content
        self.embedding, self.context, self.document_id = embedding, context, document_id


# ---------------------------------------------------------------------------------------------
# f-2 persistence
@pytest.mark.parametrize("dtype", ["bf16", "f32"])
def test_persisted_collection_reloads_bit_identical(frb, tmp_path, monkeypatch, dtype):
    monkeypatch.setenv("B200_CHILD_DTYPE", dtype)
    monkeypatch.setenv("CHROMA_CHILD_PERSIST_DIR", str(tmp_path))
    frb.reset_registry()
    n = 3000
    corpus = make_corpus(n, 384, seed=11, dup_pairs=[(7, 2500)])
    queries = make_queries(9, corpus, seed=12)
    queries[0] = corpus[7]
    store = frb.get_child_vector_store(collection="children_persist")
    kids = [_Child(10_000 + i, i // 3, f"snippet {i}", corpus[i].tolist(), context="ctx" if i % 5 == 0 else None)
            for i in range(n)]
    for lo in range(0, n, 1000):  # three ingests -> two incremental appends
        assert store.upsert_children(kids[lo:lo + 1000]) is True
    # overwrite in place, delete, re-add (appends at the end): every kind of dirty row
    kids[5].embedding = corpus[6].tolist()
    store.upsert_children([kids[5]])
    store.col.delete(ids=[str(10_000 + i) for i in (100, 101, 2999)])
    store.upsert_children([kids[100]])
    assert store.count() == n - 2
    before = [store.search(q, top_k=10) for q in queries]
    d = os.path.join(str(tmp_path), "children_persist.b200")
    meta = frb.B200Collection.committed_state(d)
    assert meta["rows"] == n + 1 and meta["dtype"] == dtype and meta["dim"] == 384 and meta["space"] == "cosine"
    assert os.path.getsize(os.path.join(d, "rows.bin")) == (n + 1) * 384 * (2 if dtype == "bf16" else 4)
    assert os.path.getsize(os.path.join(d, "keys.bin")) == (n + 1) * 8
    assert not os.path.exists(os.path.join(d, "patch.journal"))

    frb.reset_registry()  # "restart": the GPU index is gone, the next store object reloads the shard files
    store2 = frb.get_child_vector_store(collection="children_persist")
    assert store2.count() == n - 2
    after = [store2.search(q, top_k=10) for q in queries]
    assert after == before  # ids, payloads and scores, bit for bit
    assert after[0][0]["child_id"] == "10007" and after[0][1]["child_id"] == "12500"  # tie in insertion order
    # the reload dropped the two deleted rows and rewrote the files to match
    meta = frb.B200Collection.committed_state(d)
    assert meta["rows"] == n - 2
    assert os.path.getsize(os.path.join(d, meta["keys_file"])) == (n - 2) * 8
    assert sorted(f for f in os.listdir(d) if f.endswith(".bin")) == sorted([meta["rows_file"], meta["keys_file"]])
    # the reloaded collection keeps working: upsert after reload, restart again
    store2.upsert_children([_Child(77, 1, "late arrival", corpus[7].tolist())])
    frb.reset_registry()
    store3 = frb.get_child_vector_store(collection="children_persist")
    hits = store3.search(corpus[7], top_k=3)
    assert [h["child_id"] for h in hits] == ["10007", "12500", "77"]
    assert hits[2]["payload"] == {"parent_id": "1", "snippet": "late arrival"}
    frb.reset_registry()


def test_autopersist_off_and_explicit_flush(frb, tmp_path, monkeypatch):
    monkeypatch.setenv("B200_CHILD_AUTOPERSIST", "0")
    frb.reset_registry()
    client = frb.PersistentClient(path=str(tmp_path))
    col = client.get_or_create_collection("c", metadata={"hnsw:space": "l2"}, dtype="f32")
    vecs = make_corpus(50, 64, seed=3)
    col.upsert(ids=[f"id{i}" for i in range(50)], embeddings=vecs, metadatas=[{"i": i} for i in range(50)])
    assert not os.path.exists(os.path.join(str(tmp_path), "c.b200", "meta.json"))
    r0 = col.query(query_embeddings=vecs[:3], n_results=5)
    col.persist()
    frb.reset_registry()
    col2 = frb.PersistentClient(path=str(tmp_path)).get_collection("c")
    assert col2.space == "l2" and col2.dtype == "f32" and col2.count() == 50
    r1 = col2.query(query_embeddings=vecs[:3], n_results=5)
    assert r1["ids"] == r0["ids"] and r1["distances"] == r0["distances"] and r1["metadatas"] == r0["metadatas"]
    # non-numeric ids got synthetic negative keys; new ones keep counting down after the reload
    col2.upsert(ids=["another"], embeddings=vecs[:1], metadatas=[{"i": -1}])
    assert col2.key_of("another") == -52
    frb.PersistentClient(path=str(tmp_path)).delete_collection("c")
    assert not os.path.exists(os.path.join(str(tmp_path), "c.b200"))
    frb.reset_registry()


def test_raw_export_import_c_abi(frb):
    corpus = make_corpus(2000, 384, seed=21)
    a = frb.ShardIndex(dim=384, space="cosine", dtype="bf16")
    a.upsert(corpus, np.arange(2000) + 5)
    a.delete([6, 7])
    rows, keys = a.export_raw(0, 2000)
    assert rows.shape == (2000, 768) and keys[0] == 5 and keys[1] == np.iinfo(np.int64).min
    b = frb.ShardIndex(dim=384, space="cosine", dtype="bf16")
    b.import_raw(rows[:1200], keys[:1200])
    b.import_raw(rows[1200:], keys[1200:])
    assert b.count() == 1998 and b.rows() == 2000
    q = make_queries(5, corpus, seed=22)
    da, ka = a.search(q, 10)
    db, kb = b.search(q, 10)
    np.testing.assert_array_equal(ka, kb)
    np.testing.assert_array_equal(da, db)
    np.testing.assert_array_equal(b.lookup_rows([5, 6, 2004, 99999]), [0, -1, 1999, -1])
    with pytest.raises(ValueError):
        b.import_raw(rows[:, :100], keys)
    a.close()
    b.close()


def test_document_level_helpers(frb, tmp_path, monkeypatch):
    """count_for_document / delete_by_document_id are probed with hasattr by api_server.py:230-231,
    267-270; the reference's ChromaChildStore has neither."""
    monkeypatch.setenv("CHROMA_CHILD_PERSIST_DIR", str(tmp_path))
    frb.reset_registry()
    store = frb.get_child_vector_store(collection="children_docs")
    vecs = make_corpus(12, 384, seed=5)
    kids = [_Child(i + 1, 1, f"t{i}", vecs[i].tolist(), document_id="cn22-23" if i < 8 else "other") for i in range(12)]
    store.upsert_children(kids)
    assert store.count_for_document("cn22-23") == 8 and store.count_for_document("other") == 4
    assert store.count_for_document("missing") == 0
    assert store.delete_by_document_id("cn22-23") == 8
    assert store.count() == 4 and store.count_for_document("cn22-23") == 0
    assert {h["child_id"] for h in store.search(vecs[0], top_k=10)} == {"9", "10", "11", "12"}
    frb.reset_registry()


# ---------------------------------------------------------------------------------------------
# f-3 MaxSim
def _maxsim_oracle(dist, keys, shift, k_out):
    per_token = [[(str(int(k) >> shift), float(d)) for d, k in zip(dr, kr) if int(k) != -1] for dr, kr in zip(dist, keys)]
    return ofusion.maxsim_aggregate(per_token, k_out)


@pytest.mark.parametrize("T,kp,k_out", [(1, 1, 1), (4, 10, 24), (14, 10, 24), (16, 64, 5), (3, 7, 100)])
def test_maxsim_kernel_bit_exact(frb, T, kp, k_out):
    """K6 against the plain-Python restatement of multivector_store.py:155-176: identical child order
    (ties in first-seen order) and bit-identical fp64 scores."""
    rng = np.random.default_rng(T * 100 + kp)
    B = 3
    n_children = max(2, (T * kp) // 3)  # plenty of repeats inside and across tokens
    dist = rng.random((B, T, kp)).astype(np.float32)
    dist[:, :, 1::3] = dist[:, :, 0::3][:, :, : dist[:, :, 1::3].shape[2]]  # exact ties
    child = rng.integers(0, n_children, size=(B, T, kp))
    tok = rng.integers(0, 1 << 16, size=(B, T, kp))
    keys = (child.astype(np.int64) << 16) | tok
    keys[0, -1, kp // 2:] = -1  # a short list
    sc, grp = frb.maxsim_aggregate_host(dist, keys, 16, k_out)
    for b in range(B):
        want = _maxsim_oracle(dist[b], keys[b], 16, k_out)
        got_ids = [str(g) for g in grp[b] if g != -1]
        assert got_ids == [c for c, _ in want]
        assert [float(s) for s, g in zip(sc[b], grp[b]) if g != -1] == [s for _, s in want]  # bit-exact
    # device entry point gives the same
    sc_d, grp_d = frb.maxsim_aggregate_device(torch.tensor(dist).cuda(), torch.tensor(keys).cuda(), 16, k_out)
    np.testing.assert_array_equal(sc_d.cpu().numpy(), sc)
    np.testing.assert_array_equal(grp_d.cpu().numpy(), grp)


def test_multivector_store_end_to_end(frb, tmp_path, monkeypatch):
    """B200MultiVectorChildStore (mirror of multivector_store.py:27-187): one batched scan for all query
    tokens + K6 equals the reference's per-token loop restated by the oracle, and survives a restart."""
    monkeypatch.setenv("CHROMA_CHILD_PERSIST_DIR", str(tmp_path))
    monkeypatch.setenv("MULTIVECTOR_TOPK_PER_TOKEN", "10")
    frb.reset_registry()
    dim = 384
    table = {}

    def embedder(text, max_tokens):
        out = []
        for w in text.split()[:max_tokens]:
            if w not in table:
                table[w] = np.random.default_rng(abs(hash(w)) % (2 ** 32)).standard_normal(dim).astype(np.float32)
            out.append(table[w])
        return out

    words = [f"w{i}" for i in range(400)]
    rng = np.random.default_rng(1)
    kids = [_Child(5000 + c, c // 4, " ".join(rng.choice(words, size=int(rng.integers(5, 30)))), None) for c in range(120)]
    store = frb.B200MultiVectorChildStore(token_embedder=embedder)
    assert store.upsert_child_tokens(kids) is True
    n_tokens = sum(len(k.content.split()) for k in kids)
    assert store.col.count() == n_tokens
    query = " ".join(kids[17].content.split()[:6] + kids[40].content.split()[:4])
    hits = store.search_aggregate(query, top_k_children=24)
    assert len(hits) > 2 and set(hits[0]) == {"score", "child_id", "payload"}
    assert hits[0]["child_id"] in ("5017", "5040")
    assert hits[0]["payload"]["snippet"] in (kids[17].content, kids[40].content)
    # the reference's loop over the same per-token lists (same kernel, same distances) -> identical result
    qv = store._embed_tokens(query, store.query_max_tokens)
    res = store.col.query(query_embeddings=qv, n_results=10, include=["metadatas", "distances"])
    per_token = [[(m["child_id"], d) for m, d in zip(ms, ds)] for ms, ds in zip(res["metadatas"], res["distances"])]
    want = ofusion.maxsim_aggregate(per_token, 24)
    assert [h["child_id"] for h in hits] == [c for c, _ in want]
    assert [h["score"] for h in hits] == [s for _, s in want]
    # against the CPU exact scan of the token vectors (fp32): same children, scores to bf16 storage tolerance
    all_vecs = np.stack([table[w] for k in kids for w in k.content.split()])
    owner = [str(k.child_id) for k in kids for _ in k.content.split()]
    d, r = ox.exact_topk(np.asarray(qv), all_vecs, 10, "cosine", "f32")
    want2 = ofusion.maxsim_aggregate([[(owner[int(x)], dd) for x, dd in zip(rr, dr)] for rr, dr in zip(r, d)], 24)
    assert [h["child_id"] for h in hits][:3] == [c for c, _ in want2][:3]
    np.testing.assert_allclose([h["score"] for h in hits][:3], [s for _, s in want2][:3], rtol=2e-3)
    # restart: ordinals are rebuilt from the persisted payload
    frb.reset_registry()
    store2 = frb.B200MultiVectorChildStore(token_embedder=embedder)
    assert store2.search_aggregate(query, top_k_children=24) == hits
    # soft-disabled without a model, like the reference (multivector_store.py:80-86)
    store3 = frb.B200MultiVectorChildStore(collection="other_tokens")
    assert store3.upsert_child_tokens(kids) is True and store3.search_aggregate("w1 w2") == []
    frb.reset_registry()


# ---------------------------------------------------------------------------------------------
# f-1 BM25 + dense merge behind the dual-encoder ensemble (cfg3 in miniature)
class _HashEmbedder:
    """Stands in for local_embedder.SentenceTransformerWrapper: ``encode`` returns (1, d) like the wrapper."""

    def __init__(self, seed, table):
        self.seed, self.table = seed, table

    def encode(self, text, convert_to_numpy=True):
        return self.table[(self.seed, text)][None, :]


def test_hybrid_retrieval_matches_reference_pipeline(frb, tmp_path, monkeypatch):
    from financial_rag_b200.hybrid import retrieve_children_hybrid
    from oracle import bm25 as obm

    monkeypatch.setenv("CHROMA_CHILD_PERSIST_DIR", str(tmp_path))
    frb.reset_registry()
    rng = np.random.default_rng(3)
    n, dim = 400, 384
    vocab = [f"term{i}" for i in range(60)]
    texts = [" ".join(rng.choice(vocab, size=int(rng.integers(8, 40)))) for _ in range(n)]
    queries = ["term1 term2 revenue", "term2 term7 term9 growth", "term1 fiscal"]
    table, ensemble, corpora = {}, [], {}
    for seed, name in ((0, "BAAI/bge-small-en-v1.5"), (1, "thenlper/gte-small")):
        base = make_corpus(n, dim, seed=100 + seed)
        corpora[name] = base
        for qi, q in enumerate(queries):  # each query variant sits near a few children, differently per encoder
            table[(seed, q)] = (base[10 * qi + seed] + base[50 + qi] + 0.3 * make_corpus(1, dim, seed=7 + qi)[0]).astype(np.float32)
        store = frb.get_child_vector_store(collection=f"children_{seed}")
        store.upsert_children([_Child(9000 + i, i // 3, texts[i], base[i].tolist(), context="extra ctx" if i % 7 == 0 else None)
                               for i in range(n)])
        ensemble.append({"name": name, "embedder": _HashEmbedder(seed, table), "vec": store})
    chunks, child_parent, out_q = retrieve_children_hybrid(queries, ensemble, max_children=24)
    assert out_q == queries and len(chunks) == 24
    assert set(chunks[0]) == {"chunk_id", "chunk_text", "text", "retrieval_score", "retrieval_method", "child_id"}

    # the reference's pipeline restated by the oracle over the same per-(variant, encoder) hit lists
    ranked, payloads = [], {}
    for q in queries:
        for m in ensemble:
            res = m["vec"].search(m["embedder"].encode(q)[0].astype(float).tolist(), top_k=24)
            ranked.append([h["child_id"] for h in res])
            for h in res:
                payloads.setdefault(h["child_id"], h["payload"])
    dense = dict(ofusion.rrf_fuse(ranked, 60))
    docs = {c: ((p["snippet"] + "\n" + p["context"]).strip() if p.get("context") else p["snippet"]) for c, p in payloads.items()}
    want = obm.hybrid_merge(dense, docs, queries, 24)
    assert [c["child_id"] for c in chunks] == [c for c, _ in want]
    assert [c["retrieval_score"] for c in chunks] == [s for _, s in want]  # fp64, bit-exact
    assert chunks[0]["chunk_id"] == "child_" + chunks[0]["child_id"] and chunks[0]["text"] == docs[chunks[0]["child_id"]]
    assert child_parent[chunks[0]["child_id"]] == (int(chunks[0]["child_id"]) - 9000) // 3
    # dead-in-the-reference "avg" fusion stays available and agrees with its oracle too
    chunks_avg, _, _ = retrieve_children_hybrid(queries, ensemble, max_children=10, fusion="avg")
    ranked_sc = []
    for q in queries:
        for m in ensemble:
            res = m["vec"].search(m["embedder"].encode(q)[0].astype(float).tolist(), top_k=10)
            ranked_sc.append([(h["child_id"], h["score"]) for h in res])
    docs10 = {}
    for q in queries:
        for m in ensemble:
            for h in m["vec"].search(m["embedder"].encode(q)[0].astype(float).tolist(), top_k=10):
                p = h["payload"]
                docs10.setdefault(h["child_id"], (p["snippet"] + "\n" + p["context"]).strip() if p.get("context") else p["snippet"])
    want_avg = obm.hybrid_merge(dict(ofusion.avg_fuse(ranked_sc)), docs10, queries, 10)
    assert [c["child_id"] for c in chunks_avg] == [c for c, _ in want_avg]
    # (the batched scan and the one-by-one scans are different kernels: scores agree to fp32 summation noise)
    np.testing.assert_allclose([c["retrieval_score"] for c in chunks_avg], [s for _, s in want_avg], rtol=1e-5)
    frb.reset_registry()


def test_cfg3_dual_encoder_ensemble_on_device(frb):
    """cfg3 in miniature: two collections, per-collection top-50 for a query batch, RRF(k=60) to top-10 with
    nothing leaving the GPU between the scans and the fusion; ids and fp64 scores equal the oracle pipeline."""
    n, B, kp = 60000, 33, 50
    dev = torch.device("cuda", 0)
    corp = [make_corpus(n, 384, seed=s) for s in (61, 62)]
    corp[1][:2000] = corp[0][:2000] + 0.05 * make_corpus(2000, 384, seed=63)  # the encoders agree on some children
    qs = [make_queries(B, c, seed=70) for c in corp]
    cols = []
    for c in corp:
        ix = frb.ShardIndex(dim=384, space="cosine", dtype="bf16")
        ix.upsert(c, np.arange(n, dtype=np.int64))
        cols.append(ix)
    keys = torch.empty((2, B, kp), dtype=torch.int64, device=dev)
    for i, ix in enumerate(cols):
        ix.search_device(torch.tensor(qs[i]).to(dev), kp, None, keys[i])
    sc, fused = frb.rrf_fuse_device(keys, 60, 10)
    sc, fused = sc.cpu().numpy(), fused.cpu().numpy()
    ref_rows = []
    for i, ix in enumerate(cols):  # oracle: exact fp32 scan of what each collection stores
        _, r = ox.exact_topk(ox.prepare_queries(qs[i], "cosine"), ix.get_rows(0, n)[0], kp, "cosine", "f32", prepared=True)
        ref_rows.append(r)
    for b in range(B):
        lists = [[str(int(x)) for x in ref_rows[i][b]] for i in range(2)]
        want = ofusion.rrf_fuse(lists, 60, 10)
        assert [str(int(x)) for x in fused[b]] == [c for c, _ in want], f"query {b}"
        assert [float(s) for s in sc[b]] == [s for _, s in want]
    for ix in cols:
        ix.close()


def test_concurrent_search_and_upsert_threads(frb, tmp_path, monkeypatch):
    """The reference serves requests from Flask threads + an executor thread + a daemon ingest thread
    (api_server.py:851-866, 1342-1371), each constructing its own store object: concurrent search and
    upsert_children on one collection must stay consistent (threading)."""
    import threading

    monkeypatch.setenv("CHROMA_CHILD_PERSIST_DIR", str(tmp_path))
    monkeypatch.setenv("B200_CHILD_AUTOPERSIST", "0")
    frb.reset_registry()
    n0, n_add = 4000, 2000
    corpus = make_corpus(n0 + n_add, 384, seed=31)
    store = frb.get_child_vector_store(collection="children_threads")
    store.upsert_children([_Child(i + 1, 0, f"t{i}", corpus[i].tolist()) for i in range(n0)])
    errors, done = [], threading.Event()

    def searcher(tid):
        try:
            s = frb.get_child_vector_store(collection="children_threads")  # a store object per request
            rng = np.random.default_rng(tid)
            while not done.is_set():
                j = int(rng.integers(0, n0))
                hits = s.search(corpus[j], top_k=5)
                assert hits[0]["child_id"] == str(j + 1) and hits[0]["score"] > 0.99, (j, hits[0])
                multi = s.search_batch(corpus[j:j + 3], top_k=5)
                assert [m[0]["child_id"] for m in multi] == [str(j + 1 + t) for t in range(len(multi))]
        except Exception as e:  # noqa: BLE001
            errors.append(repr(e))

    def writer():
        try:
            s = frb.get_child_vector_store(collection="children_threads")
            for lo in range(n0, n0 + n_add, 100):
                s.upsert_children([_Child(i + 1, 0, f"t{i}", corpus[i].tolist()) for i in range(lo, lo + 100)])
        except Exception as e:  # noqa: BLE001
            errors.append(repr(e))

    threads = [threading.Thread(target=searcher, args=(t,)) for t in range(4)]
    w = threading.Thread(target=writer)
    for t in threads + [w]:
        t.start()
    w.join()
    done.set()
    for t in threads:
        t.join()
    assert not errors, errors[:3]
    assert store.count() == n0 + n_add
    assert store.search(corpus[n0 + n_add - 1], top_k=1)[0]["child_id"] == str(n0 + n_add)
    frb.reset_registry()


def test_migrate_chroma_wal_replay(frb, golden, tmp_path, monkeypatch):
    """A Chroma persist directory with the schema of the reference's .chroma_children/chroma.sqlite3 (built here
    from the golden fixture: same 18 WAL rows, same blobs) is replayed into B200 collections; the migrated
    collections answer the known-answer query of SURVEY.md 8c (A1 A2 A3 B1 B2 B3 C1 C2 C3, insertion order)."""
    import sqlite3

    from financial_rag_b200.migrate_chroma import replay_wal

    src = tmp_path / "chroma_src"
    src.mkdir()
    con = sqlite3.connect(str(src / "chroma.sqlite3"))
    con.executescript("""
        create table collections (id text primary key, name text not null, dimension integer, database_id text not null,
                                  config_json_str text);
        create table collection_metadata (collection_id text, key text, str_value text, int_value integer,
                                          float_value real, bool_value integer);
        create table segments (id text primary key, type text not null, scope text not null, collection text not null);
        create table embeddings (id integer primary key, segment_id text not null, embedding_id text not null,
                                 seq_id blob not null);
        create table embeddings_queue (seq_id integer primary key, operation integer not null, topic text not null,
                                       id text not null, vector blob, encoding text, metadata text);
    """)
    uuid_of = {}
    for i, (name, cfg) in enumerate(golden["raw"]["collections"].items()):
        uuid_of[name] = f"0000000{i}-aaaa-bbbb-cccc-00000000000{i}"
        con.execute("insert into collections values (?,?,?,?,?)",
                    (uuid_of[name], name, cfg["dimension"], "db", json.dumps(cfg["config"])))
        con.execute("insert into segments values (?,?,?,?)", (f"seg{i}", "urn:vector", "VECTOR", uuid_of[name]))
    for r in golden["raw"]["rows"]:
        con.execute("insert into embeddings_queue values (?,?,?,?,?,?,?)",
                    (r["seq_id"], r["operation"], f"persistent://default/default/{uuid_of[r['collection']]}", r["id"],
                     bytes.fromhex(r["vector_f32le_hex"]), "FLOAT32", json.dumps(r["metadata"]) if r["metadata"] else None))
    # one delete + one re-upsert at the end of the log, to exercise every operation
    first = golden["raw"]["rows"][0]
    topic = f"persistent://default/default/{uuid_of[first['collection']]}"
    con.execute("insert into embeddings_queue values (100, 3, ?, ?, null, null, null)", (topic, first["id"]))
    con.execute("insert into embeddings_queue values (101, 2, ?, ?, ?, 'FLOAT32', ?)",
                (topic, first["id"], bytes.fromhex(first["vector_f32le_hex"]), json.dumps(first["metadata"])))
    con.commit()
    con.close()

    frb.reset_registry()
    dst = frb.PersistentClient(path=str(tmp_path / "b200_dst"))
    report = replay_wal(str(src), dst)
    for name, col in golden["collections"].items():
        assert report[name]["count"] == 9, report
        got = dst.get_collection(name)
        assert got.space == "cosine"
        res = got.query(query_embeddings=[col["vectors"][1].tolist()], n_results=9, include=["metadatas", "distances"])
        want = [col["ids"][i] for i in (1, 4, 7, 0, 3, 6, 2, 5, 8)]
        if name == first["collection"]:  # the deleted + re-upserted id moved to the end of the insertion order
            want = [col["ids"][i] for i in (1, 4, 7, 3, 6, 0, 2, 5, 8)]
        assert res["ids"][0] == want
        assert res["metadatas"][0][0] == col["metadatas"][1]
    frb.reset_registry()


# ---------------------------------------------------------------------------------------------
# round-2 additions: shared ordinals, crash-safe flushes, the directory lock, score fusion, the ensemble searcher
def test_multivector_store_on_a_row_sharded_collection(frb, tmp_path, monkeypatch):
    """B200_CHILD_DEVICES also shards the token collection of the multi-vector store: same MaxSim ranking."""
    rng = np.random.default_rng(9)
    table = {}

    def embedder(text, max_tokens):
        return [table.setdefault(w, rng.standard_normal(384).astype(np.float32)) for w in text.split()[:max_tokens]]

    words = [f"w{i}" for i in range(300)]
    kids = [_Child(7000 + c, c // 4, " ".join(rng.choice(words, size=int(rng.integers(5, 25)))), None) for c in range(80)]
    query = " ".join(kids[11].content.split()[:5] + kids[30].content.split()[:3])
    answers = []
    for devices in (None, "0,0,0"):
        if devices is None:
            monkeypatch.delenv("B200_CHILD_DEVICES", raising=False)
        else:
            monkeypatch.setenv("B200_CHILD_DEVICES", devices)
        monkeypatch.setenv("CHROMA_CHILD_PERSIST_DIR", str(tmp_path / (devices or "one").replace(",", "_")))
        frb.reset_registry()
        store = frb.B200MultiVectorChildStore(token_embedder=embedder)
        store.upsert_child_tokens(kids)
        answers.append(store.search_aggregate(query, top_k_children=10))
        frb.reset_registry()
    assert [h["child_id"] for h in answers[1]] == [h["child_id"] for h in answers[0]]
    np.testing.assert_allclose([h["score"] for h in answers[1]], [h["score"] for h in answers[0]], rtol=0, atol=1e-5)
    assert answers[0][0]["child_id"] in ("7011", "7030")


def test_multivector_store_objects_share_the_collections_ordinals(frb, tmp_path, monkeypatch):
    """The reference builds a MultiVectorChildStore per request (rag_backend.py:656) and another for ingest
    (pipeline.py:25).  Two live objects that ingest different children must not hand out the same ordinal: the map
    lives in the collection."""
    monkeypatch.setenv("CHROMA_CHILD_PERSIST_DIR", str(tmp_path))
    frb.reset_registry()
    rng = np.random.default_rng(5)
    table = {}

    def embedder(text, max_tokens):
        return [table.setdefault(w, rng.standard_normal(384).astype(np.float32)) for w in text.split()[:max_tokens]]

    a = frb.B200MultiVectorChildStore(token_embedder=embedder)
    b = frb.B200MultiVectorChildStore(token_embedder=embedder)   # constructed BEFORE a's ingest: no stale state to go wrong
    a.upsert_child_tokens([_Child(1, 10, "alpha beta gamma"), _Child(2, 10, "delta epsilon")])
    b.upsert_child_tokens([_Child(3, 11, "zeta eta theta iota"), _Child(1, 10, "alpha beta gamma")])
    assert a.col is b.col and a.col.count() == 3 + 2 + 4
    assert [a.col.group_ordinal(str(c), create=False) for c in (1, 2, 3)] == [0, 1, 2]
    for store in (a, b, frb.B200MultiVectorChildStore(token_embedder=embedder)):
        hits = store.search_aggregate("zeta eta", top_k_children=3)
        assert hits[0]["child_id"] == "3" and hits[0]["payload"]["snippet"] == "zeta eta theta iota"
        hits = store.search_aggregate("delta", top_k_children=3)
        assert hits[0]["child_id"] == "2" and hits[0]["payload"]["parent_id"] == "10"
    # an explicit key that belongs to another id is refused instead of silently overwriting its row and payload
    with pytest.raises(ValueError):
        a.col.upsert(ids=["99:0"], embeddings=rng.standard_normal((1, 384)), metadatas=[{}], keys=[(0 << 16) | 1])
    frb.reset_registry()
    c = frb.B200MultiVectorChildStore(token_embedder=embedder)   # restart: ordinals come back with the collection
    assert [c.col.group_ordinal(str(x), create=False) for x in (1, 2, 3)] == [0, 1, 2]
    assert c.search_aggregate("zeta eta", top_k_children=3)[0]["child_id"] == "3"
    frb.reset_registry()


def test_flush_survives_a_crash_before_and_after_the_commit(frb, tmp_path, monkeypatch):
    """persist() = append past the committed count -> journal the in-place patches -> ONE sqlite commit -> apply the
    patches.  Kill it before the commit: the reload is the old collection, bit for bit.  Kill it after: the reload
    finishes the flush.  Never new vector bits under an old payload."""
    monkeypatch.setenv("CHROMA_CHILD_PERSIST_DIR", str(tmp_path))
    monkeypatch.setenv("B200_CHILD_DTYPE", "f32")
    frb.reset_registry()
    corpus = make_corpus(300, 384, seed=31)
    store = frb.get_child_vector_store(collection="crashy")
    store.upsert_children([_Child(100 + i, 1, f"old {i}", corpus[i].tolist()) for i in range(200)])
    col = store.col
    before = store.search(corpus[250].tolist(), top_k=5)

    class Boom(RuntimeError):
        pass

    def boom(*a, **k):
        raise Boom()

    # (a) crash between the journal and the commit: child 100 is overwritten with vector 250, ten children appended
    monkeypatch.setattr(col, "_open_payload_db", boom)
    with pytest.raises(Boom):
        store.upsert_children([_Child(100, 1, "NEW 0", corpus[250].tolist())] +
                              [_Child(1000 + i, 2, f"app {i}", corpus[200 + i].tolist()) for i in range(10)])
    assert os.path.exists(os.path.join(col.directory, "patch.journal"))
    frb.reset_registry()  # "restart"
    store = frb.get_child_vector_store(collection="crashy")
    assert store.count() == 200 and not os.path.exists(os.path.join(store.col.directory, "patch.journal"))
    assert store.search(corpus[250].tolist(), top_k=5) == before
    assert store.search(corpus[0].tolist(), top_k=1)[0]["payload"]["snippet"] == "old 0"
    # (b) crash right after the commit, before the patches reach rows.bin
    col = store.col
    monkeypatch.setattr(col, "_apply_patches", boom)
    with pytest.raises(Boom):
        store.upsert_children([_Child(100, 1, "NEW 0", corpus[250].tolist())] +
                              [_Child(1000 + i, 2, f"app {i}", corpus[200 + i].tolist()) for i in range(10)])
    frb.reset_registry()
    store = frb.get_child_vector_store(collection="crashy")
    assert store.count() == 210
    top = store.search(corpus[250].tolist(), top_k=1)[0]
    assert top["child_id"] == "100" and top["payload"]["snippet"] == "NEW 0" and top["score"] > 0.9999
    assert store.search(corpus[205].tolist(), top_k=1)[0]["child_id"] == "1005"
    # (c) a damaged directory is reported, not guessed at
    frb.reset_registry()
    rows_p = os.path.join(str(tmp_path), "crashy.b200", "rows.bin")
    with open(rows_p, "r+b") as f:
        f.truncate(os.path.getsize(rows_p) - 1536 * 3)
    with pytest.raises(RuntimeError, match="damaged"):
        frb.get_child_vector_store(collection="crashy")
    frb.reset_registry()


def test_collection_directory_is_single_writer(frb, tmp_path, monkeypatch):
    """A second PROCESS on the same persist directory (ingest_all.py next to the API server) is refused with a clear
    error; it used to truncate the rows the first one had appended."""
    import subprocess

    monkeypatch.setenv("CHROMA_CHILD_PERSIST_DIR", str(tmp_path))
    frb.reset_registry()
    store = frb.get_child_vector_store(collection="locked")
    store.upsert_children([_Child(1, 1, "x", make_corpus(1, 384, seed=1)[0].tolist())])
    code = ("import os, sys; sys.path.insert(0, %r); import financial_rag_b200 as f\n"
            "try:\n    f.get_child_vector_store(collection='locked').count(); print('OPENED')\n"
            "except RuntimeError as e:\n    print('REFUSED', e)\n" % os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=dict(os.environ), timeout=300).stdout
    assert "REFUSED" in out and "another process" in out, out
    frb.reset_registry()  # closing releases the lock
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=dict(os.environ), timeout=300).stdout
    assert "OPENED" in out, out


@pytest.mark.parametrize("L,B,kp,k_out", [(1, 1, 1, 1), (2, 3, 5, 10), (6, 4, 30, 24), (3, 2, 50, 200)])
def test_score_fuse_kernel_bit_exact(frb, L, B, kp, k_out):
    """K5b == rag_backend.py:732-754 (per-list min-max normalised scores, mean over lists) in fp64, bit for bit:
    constant lists, empty lists, short lists, keys shared between lists."""
    rng = np.random.default_rng(L * 100 + kp)
    dist = rng.random((L, B, kp)).astype(np.float32)
    keys = rng.integers(0, max(2, kp * 2), size=(L, B, kp)).astype(np.int64)
    for l in range(L):
        for b in range(B):
            keys[l, b] = rng.permutation(max(2, kp * 2))[:kp]   # ids are unique inside one list
    if L > 1:
        dist[1, 0, :] = 0.25         # a constant list contributes 0 for every member
        keys[0, B - 1, kp // 2:] = -1  # a short list
    if L > 2:
        keys[2, 0, :] = -1           # an empty list still counts in the divisor
    sc, fused = frb.score_fuse_host(dist, keys, k_out)
    for b in range(B):
        lists = []
        for l in range(L):
            n = int((keys[l, b] != -1).sum())
            lists.append([(str(int(k)), 1.0 - float(d)) for k, d in zip(keys[l, b, :n], dist[l, b, :n])])
        want = ofusion.avg_fuse(lists, k_out)
        got_ids = [str(int(k)) for k in fused[b] if k != -1]
        assert got_ids == [c for c, _ in want], f"query {b}"
        assert [float(s) for s, k in zip(sc[b], fused[b]) if k != -1] == [s for _, s in want]


@pytest.mark.parametrize("fusion", ["rrf", "avg"])
@pytest.mark.parametrize("sharded", [False, True])
def test_ensemble_searcher_device_resident(frb, fusion, sharded):
    """EnsembleSearcher = cfg3 as a product call: every collection scanned, the lists fused on the device, the same
    answer whether the collections are one shard or row-sharded groups, equal to the oracle pipeline."""
    n, B, kp = 40000, 20, 50
    corp = [make_corpus(n, 384, seed=s) for s in (81, 82)]
    corp[1][:1500] = corp[0][:1500] + 0.05 * make_corpus(1500, 384, seed=83)
    qs = [make_queries(B, c, seed=90) for c in corp]
    cols = []
    for c in corp:
        ix = frb.ShardGroup(dim=384, dtype="bf16", devices=[0, 0, 0]) if sharded else frb.ShardIndex(dim=384, dtype="bf16")
        ix.upsert(c, np.arange(n, dtype=np.int64))
        cols.append(ix)
    ens = frb.EnsembleSearcher(cols, k_each=kp, k_rrf=60, k_out=10, fusion=fusion)
    sc, fused = ens.search(qs)
    lists_d, lists_k = [], []
    for ix, q in zip(cols, qs):
        d, kk = ix.search(q, kp)
        lists_d.append(d)
        lists_k.append(kk)
    for b in range(B):
        if fusion == "rrf":
            want = ofusion.rrf_fuse([[str(int(x)) for x in lk[b]] for lk in lists_k], 60, 10)
        else:
            want = ofusion.avg_fuse([[(str(int(x)), 1.0 - float(dd)) for x, dd in zip(lk[b], ld[b])]
                                     for lk, ld in zip(lists_k, lists_d)], 10)
        assert [str(int(x)) for x in fused[b]] == [c for c, _ in want], f"query {b}"
        assert [float(s) for s in sc[b]] == [s for _, s in want]
    for ix in cols:
        ix.close()
