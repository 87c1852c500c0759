"""GPU parity tests of the row-sharded collection (fr_group / ShardGroup).

The bar: a collection spread over W shards answers exactly like the same collection on one GPU -- ids, order (ties in
global insertion order) and distances -- through the C ABI and through the reference-facing store API.

On a one-GPU box the W shards all sit on device 0 and bring their lists together without NCCL (FR_XCHG_PEER: the shards'
last kernels store straight into the merging buffer; FR_XCHG_COPY: peer copies); the tests marked ``multi`` need W
distinct GPUs and cover NCCL, peer stores over NVLink and peer copies (they skip on smaller boxes; bench.py's pre-flight
and ``gpurun --gpus N`` runs cover them).
"""
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")

from oracle import exact_scan as ox  # noqa: E402

from helpers import assert_matches_oracle, keys_to_rows, make_corpus, make_queries  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BGE = "children_baai_bge_small_en_v1_5"
KEY_BASE = 5000


@pytest.fixture(scope="module")
def frb():
    if not torch.cuda.is_available():
        pytest.fail("GPU tests need a CUDA device (there is no CPU fallback to test)")
    import financial_rag_b200 as f

    return f


def _devices(world, distinct):
    if distinct:
        if torch.cuda.device_count() < world:
            pytest.skip(f"needs {world} GPUs, box has {torch.cuda.device_count()}")
        return list(range(world))
    return [0] * world


def _fill(ix, corpus, key_base=KEY_BASE, chunk=None):
    n = corpus.shape[0]
    keys = np.arange(n, dtype=np.int64) + key_base
    step = chunk or max(n, 1)
    for lo in range(0, n, step):
        ix.upsert(corpus[lo:lo + step], keys[lo:lo + step])
    return keys


def _same_answer(a, b, label):
    (da, ka), (db, kb) = a, b
    assert (ka == kb).all(), f"{label}: keys differ at {np.argwhere(ka != kb)[:3].tolist()}"
    fin = np.isfinite(db)
    assert (np.isfinite(da) == fin).all()
    # the shards may take different scan kernels than the single index (their row counts differ): last-ulp noise only
    assert np.abs(da[fin] - db[fin]).max(initial=0.0) <= 2e-6, f"{label}: distances differ"


CASES = [
    # rows, world, batch, k, space, dtype
    (9, 4, 3, 10, "cosine", "f32"),          # fewer rows than k, shards with 2-3 rows
    (3, 8, 1, 5, "cosine", "bf16"),          # empty shards
    (5000, 2, 5, 10, "cosine", "bf16"),
    (5000, 3, 70, 10, "l2", "f32"),
    (20000, 8, 33, 32, "ip", "bf16"),
    (30011, 4, 300, 50, "cosine", "bf16"),   # tensor-core path on every shard, k' = 128
    (30011, 3, 130, 100, "cosine", "f32"),
]


@pytest.mark.parametrize("exchange", ["peer", "copy"])
@pytest.mark.parametrize("n,world,B,k,space,dtype", CASES)
def test_group_equals_single_index_and_oracle(frb, n, world, B, k, space, dtype, exchange):
    corpus = make_corpus(n, seed=n + world, dup_pairs=[(0, n - 1), (1, n // 2)] if n > 4 else [(0, 2)])
    q = make_queries(B, corpus, seed=7)
    q[0] = corpus[0]  # exact ties across shards: rows 0 and n-1 are byte-identical
    one = frb.ShardIndex(dim=384, space=space, dtype=dtype)
    grp = frb.ShardGroup(dim=384, space=space, dtype=dtype, devices=_devices(world, False), exchange=exchange)
    assert grp.exchange == exchange and grp.world == world
    _fill(one, corpus, chunk=1777)
    _fill(grp, corpus, chunk=1777)
    assert grp.count() == one.count() == n and grp.rows() == n
    # cyclic placement: shard s holds rows s, s + W, ...
    for s in range(world):
        assert grp.shard(s).rows() == len(range(s, n, world))
    got, want = grp.search(q, k), one.search(q, k)
    _same_answer(got, want, f"W={world}")
    stored, stored_keys = grp.get_rows(0, n)
    assert (stored_keys == np.arange(n) + KEY_BASE).all()
    assert (stored == one.get_rows(0, n)[0]).all()
    assert_matches_oracle(got[0], keys_to_rows(got[1], KEY_BASE), q, corpus, k, space, dtype, stored=stored,
                          label=f"group W={world}")
    if n > 4:  # the tie: row 0 before its copy n-1, whatever shards they sit on
        assert got[1][0, 0] == KEY_BASE and got[1][0, 1] == KEY_BASE + n - 1
    grp.close()
    one.close()


def test_group_upsert_delete_overwrite_semantics(frb):
    rng = np.random.default_rng(3)
    corpus = rng.standard_normal((1000, 384), dtype=np.float32)
    one = frb.ShardIndex(dim=384, dtype="f32")
    grp = frb.ShardGroup(dim=384, dtype="f32", devices=[0, 0, 0])
    for ix in (one, grp):
        keys = _fill(ix, corpus, chunk=333)
        # overwrite in place (keeps the insertion position), a key repeated inside one call: last one wins
        ix.upsert(np.stack([corpus[5], corpus[6], corpus[7]]), np.array([keys[10], keys[11], keys[11]]))
        assert ix.delete(np.array([keys[0], keys[500], 123456789])) == 2
        assert ix.count() == 998 and ix.rows() == 1000
        # appended after the delete: new rows at the end
        ix.upsert(corpus[:3] * 1.5, np.array([9_000_001, 9_000_002, keys[0]]))
        assert ix.count() == 1001 and ix.rows() == 1003
    q = np.concatenate([corpus[5:8], corpus[:1], corpus[500:501], rng.standard_normal((4, 384), dtype=np.float32)])
    _same_answer(grp.search(q, 10), one.search(q, 10), "after mutations")
    d, kk = grp.search(q, 10)
    assert kk[0, 0] == KEY_BASE + 5 and kk[0, 1] == KEY_BASE + 10      # row 10 now holds vector 5: tie in row order
    assert kk[2, :2].tolist() == [KEY_BASE + 7, KEY_BASE + 11]          # last writer of the repeated key won
    assert kk[3, 0] == 9_000_001                                        # row 0 is gone; its vector came back under a new key
    assert (grp.lookup_rows(np.array([KEY_BASE + 1, KEY_BASE + 500, 9_000_002, KEY_BASE])) == [1, -1, 1001, 1002]).all()
    with pytest.raises(Exception):
        grp.upsert(corpus[:1], np.array([-1]))  # FR_KEY_NONE is the padding sentinel
    grp.close()
    one.close()


def test_group_shard_files_are_world_size_independent(frb):
    corpus = make_corpus(4001, seed=11)
    q = make_queries(9, corpus, seed=12)
    src = frb.ShardGroup(dim=384, dtype="bf16", devices=[0, 0, 0])
    keys = _fill(src, corpus)
    src.delete(keys[100:110])
    want = src.search(q, 10)
    rows, rkeys = src.export_raw(0, src.rows())
    assert (rkeys[100:110] == np.iinfo(np.int64).min).all()
    for world in (1, 2, 5):
        dst = frb.ShardGroup(dim=384, dtype="bf16", devices=[0] * world)
        dst.import_raw(rows[:1500], rkeys[:1500])
        dst.import_raw(rows[1500:], rkeys[1500:])
        assert dst.rows() == 4001 and dst.count() == 3991
        _same_answer(dst.search(q, 10), want, f"reloaded at W={world}")
        assert dst.lookup_rows(keys[[0, 105, 4000]]).tolist() == [0, -1, 4000]
        dst.close()
    one = frb.ShardIndex(dim=384, dtype="bf16")
    one.import_raw(rows, rkeys)
    _same_answer(one.search(q, 10), want, "reloaded into one shard")
    one.close()
    src.close()


def test_group_device_api_and_bulk_load(frb):
    """The bench's way in: rows appended on each shard's device, adopt_rows, device-resident queries."""
    n, world, k = 40000, 4, 10
    dev = torch.device("cuda", 0)
    g = torch.Generator(device=dev).manual_seed(5)
    corpus = torch.randn((n, 384), generator=g, device=dev)
    grp = frb.ShardGroup(dim=384, dtype="bf16", devices=[0] * world, reserve_rows=n)
    one = frb.ShardIndex(dim=384, dtype="bf16")
    one.append_device(corpus, None, first_key=0)
    for s in range(world):
        rows = torch.arange(s, n, world, device=dev)
        grp.shard(s).append_device(corpus[rows].contiguous(), rows.contiguous())
    with pytest.raises(Exception):
        grp.adopt_rows(n + 1)
    grp.adopt_rows(n)
    assert grp.count() == n
    q = torch.nn.functional.normalize(corpus[::4001][:7] + 0.05 * torch.randn((7, 384), generator=g, device=dev))
    want_d, want_k = one.search_device(q, k)
    # bulk-loaded shards: the key map is rebuilt from the devices on demand
    assert grp.lookup_rows(np.array([0, 1, n - 1, n])).tolist() == [0, 1, n - 1, -1]
    dist, keys = grp.search_device([q] * world, k, merge_on=[0, 2])
    torch.cuda.synchronize()
    assert dist[1] is None and dist[3] is None
    for j in (0, 2):
        assert (keys[j] == want_k).all() and (dist[j] - want_d).abs().max().item() <= 2e-6
    hd, hk = grp.search(q.cpu().numpy(), k)
    assert (hk == want_k.cpu().numpy()).all()
    grp.close()
    one.close()


class _Child:
    def __init__(self, child_id, parent_id, content, embedding, context=None):
        self.child_id, self.parent_id, self.content = child_id, parent_id, content
        self.embedding, self.context = embedding, context


def _store_answers(frb, golden, monkeypatch, tmp_path, devices_env, corpus, q):
    """Ingest the reference's fixture plus a synthetic corpus through get_child_vector_store and query it."""
    if devices_env is None:
        monkeypatch.delenv("B200_CHILD_DEVICES", raising=False)
    else:
        monkeypatch.setenv("B200_CHILD_DEVICES", devices_env)
    monkeypatch.setenv("CHROMA_CHILD_PERSIST_DIR", str(tmp_path / (devices_env or "one").replace(",", "_")))
    monkeypatch.setenv("B200_CHILD_DTYPE", "f32")
    frb.reset_registry()
    col = golden["collections"][BGE]
    store = frb.get_child_vector_store(collection=BGE)
    kids = [_Child(int(c), int(m["parent_id"]), m["snippet"], v.tolist(), m.get("context"))
            for c, v, m in zip(col["ids"], col["vectors"], col["metadatas"])]
    for lo in (0, 3, 6):  # three ingests of three children, like the fixture's WAL
        assert store.upsert_children(kids[lo:lo + 3]) is True
    store.upsert_children([_Child(10_000 + i, 1, f"synthetic {i}", corpus[i]) for i in range(corpus.shape[0])])
    out = [store.search(col["vectors"][0].tolist(), top_k=10)]
    out += [frb.get_child_vector_store(collection=BGE).search(v, top_k=6) for v in q]
    n = store.count()
    # restart: the collection comes back from its shard files, on whatever devices are configured now
    frb.reset_registry()
    again = frb.get_child_vector_store(collection=BGE)
    assert again.count() == n
    assert again.search(col["vectors"][0].tolist(), top_k=10) == out[0]
    frb.reset_registry()
    return out, n


@pytest.mark.parametrize("devices_env", ["0,0", "0,0,0,0,0"])
def test_store_api_sharded_equals_unsharded(frb, golden, monkeypatch, tmp_path, devices_env):
    """get_child_vector_store(...).search over a row-sharded collection == the one-GPU answer, dict for dict."""
    rng = np.random.default_rng(21)
    corpus = rng.standard_normal((600, 384), dtype=np.float32)
    q = rng.standard_normal((4, 384), dtype=np.float32)
    want, n1 = _store_answers(frb, golden, monkeypatch, tmp_path, None, corpus, q)
    got, n2 = _store_answers(frb, golden, monkeypatch, tmp_path, devices_env, corpus, q)
    assert n1 == n2 == 609
    ids = golden["collections"][BGE]["ids"]
    assert [h["child_id"] for h in got[0]][:9] == [ids[i] for i in (0, 3, 6, 1, 4, 7, 2, 5, 8)]
    for a, b in zip(got, want):
        assert [h["child_id"] for h in a] == [h["child_id"] for h in b]
        assert [h["payload"] for h in a] == [h["payload"] for h in b]
        np.testing.assert_allclose([h["score"] for h in a], [h["score"] for h in b], rtol=0, atol=2e-6)


# ---- real multi-GPU: distinct devices, NCCL exchange ---------------------------------------------------------------
@pytest.mark.parametrize("G", [2, 4, 8])
def test_multi_gpu_store_api_equals_one_gpu(frb, golden, monkeypatch, tmp_path, G):
    if torch.cuda.device_count() < G:
        pytest.skip(f"needs {G} GPUs, box has {torch.cuda.device_count()}")
    rng = np.random.default_rng(22)
    corpus = rng.standard_normal((3000, 384), dtype=np.float32)
    q = rng.standard_normal((6, 384), dtype=np.float32)
    want, _ = _store_answers(frb, golden, monkeypatch, tmp_path, None, corpus, q)
    got, _ = _store_answers(frb, golden, monkeypatch, tmp_path, ",".join(str(i) for i in range(G)), corpus, q)
    for a, b in zip(got, want):
        assert [h["child_id"] for h in a] == [h["child_id"] for h in b]
        np.testing.assert_allclose([h["score"] for h in a], [h["score"] for h in b], rtol=0, atol=2e-6)


@pytest.mark.parametrize("G,exchange", [(2, "nccl"), (2, "copy"), (2, "peer"), (4, "nccl"), (4, "peer"), (8, "nccl"), (8, "peer")])
def test_multi_gpu_group_equals_single_index(frb, G, exchange):
    devices = _devices(G, True)
    corpus = make_corpus(60000, seed=G, dup_pairs=[(0, 59999), (1, 30000)])
    one = frb.ShardIndex(dim=384, dtype="bf16")
    grp = frb.ShardGroup(dim=384, dtype="bf16", devices=devices, exchange=exchange)
    assert grp.exchange == exchange
    _fill(one, corpus)
    _fill(grp, corpus, chunk=7001)
    for B, k in ((1, 10), (40, 10), (300, 10), (128, 100)):
        q = make_queries(B, corpus, seed=B)
        q[0] = corpus[0]
        got, want = grp.search(q, k), one.search(q, k)
        _same_answer(got, want, f"G={G} B={B} k={k}")
        assert got[1][0, 0] == KEY_BASE and got[1][0, 1] == KEY_BASE + 59999
    stored = grp.get_rows(0, 60000)[0]
    q = make_queries(64, corpus, seed=99)
    d, kk = grp.search(q, 10)
    assert_matches_oracle(d, keys_to_rows(kk, KEY_BASE), q, corpus, 10, "cosine", "bf16", stored=stored, label=f"G={G}")
    # device-resident form, every device merges
    qs = [torch.from_numpy(q).to(f"cuda:{dv}") for dv in devices]
    dist, keys = grp.search_device(qs, 10)
    for dv in devices:
        torch.cuda.synchronize(dv)
    for j in range(G):
        assert (keys[j].cpu().numpy() == kk).all()
    grp.close()
    one.close()


@pytest.mark.parametrize("world", [2, 4])
def test_multi_process_group_under_torchrun(world):
    """One process per GPU (how bench.py --gpus N runs): SPMD upserts, NCCL id over torch.distributed, every rank gets
    the one-GPU answer (tests/spmd_group_check.py asserts it on every rank and exits non-zero otherwise)."""
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs, box has {torch.cuda.device_count()}")
    env = dict(os.environ)
    env.pop("B200_CHILD_DEVICES", None)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
                        "--master-addr", "127.0.0.1", "--master-port", str(29600 + world),
                        os.path.join(ROOT, "tests", "spmd_group_check.py")],
                       capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    assert f"SPMD-OK world={world}" in r.stdout


def test_concurrent_host_searches_on_a_group(frb):
    """Several request threads on one row-sharded collection: each call holds the group's lock only while it is
    enqueued and reads its own pinned slot; every thread gets the lone call's answer."""
    import threading

    corpus = make_corpus(50000, seed=31)
    grp = frb.ShardGroup(dim=384, dtype="bf16", devices=[0, 0, 0])
    _fill(grp, corpus)
    jobs = []
    for t in range(6):
        q = make_queries((1, 5, 40, 70, 200, 300)[t], corpus, seed=40 + t)
        jobs.append((q, grp.search(q, 10)))
    errors = []

    def worker(t):
        q, (want_d, want_k) = jobs[t]
        try:
            for _ in range(15):
                d, kk = grp.search(q, 10)
                if not ((kk == want_k).all() and (d == want_d).all()):
                    errors.append(f"thread {t}: answer differs from the lone call")
                    return
        except Exception as e:  # noqa: BLE001
            errors.append(f"thread {t}: {e!r}")

    threads = [threading.Thread(target=worker, args=(t,)) for t in range(6)]
    for th in threads:
        th.start()
    for th in threads:
        th.join()
    assert not errors, errors
    grp.close()
