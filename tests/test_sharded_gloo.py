"""World-size-2 (and 3) run of the row-sharded search plumbing on CPU with the gloo backend.

The CUDA compute steps of ShardedSearcher are replaced here -- in the TEST only -- by the CPU
oracle, so what is exercised is the product's host logic: contiguous shard bounds, the packed
[2, B, k] int64 exchange buffer, one all-gather, merge order (score, shard, local row) and the
rank-0 host path with its broadcast.  Result must equal the single-shard oracle for any G.
"""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from helpers import make_corpus, make_queries

N, B, K, DIM = 5003, 6, 10, 384


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _order_bits(score_f32: np.ndarray) -> np.ndarray:
    u = score_f32.astype(np.float32).view(np.uint32).astype(np.uint64)
    return np.where(u & 0x80000000, (~u) & 0xFFFFFFFF, u | 0x80000000)


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import exact_scan as ox

    from financial_rag_b200.sharded import ShardedSearcher, shard_bounds

    corpus = make_corpus(N, DIM, seed=5, dup_pairs=[(7, 4000), (8, 2600)])
    queries = make_queries(B, corpus, seed=6)
    queries[0], queries[1] = corpus[7], corpus[8]
    lo, hi = shard_bounds(N, world, rank, align=100)
    shard = ox.prepare_corpus(corpus[lo:hi], "cosine", "bf16")
    holder = {}

    class FakeIndex:
        device = 0

    def local_search(q, b):
        qp = ox.prepare_queries(q.numpy(), "cosine")
        d, r = ox.exact_topk(qp, shard, K, "cosine", "f32", prepared=True)
        score = (np.float32(1.0) - d).astype(np.float32)
        packed = (_order_bits(score) << np.uint64(32)) | (np.uint64(0xFFFFFFFF) - r.astype(np.uint64))
        packed = np.where(r >= 0, packed, np.uint64(0))
        s = holder["s"]
        s.local[0, : b * K] = torch.from_numpy(packed.view(np.int64).reshape(-1))
        s.local[1, : b * K] = torch.from_numpy(np.where(r >= 0, r + lo, -1).reshape(-1))

    def merge(b):
        s = holder["s"]
        g = s.gathered.numpy()  # [G, 2, max_batch*K]
        for i in range(b):
            cand = []
            for gi in range(world):
                pk = g[gi, 0, i * K:(i + 1) * K].view(np.uint64)
                ky = g[gi, 1, i * K:(i + 1) * K]
                for j in range(K):
                    if pk[j] != 0:
                        cand.append((-(int(pk[j]) >> 32), gi, j, int(ky[j]), int(pk[j]) >> 32))
            cand.sort()
            for j in range(K):
                if j < len(cand):
                    ob = np.uint32(cand[j][4])
                    u = (ob & np.uint32(0x7FFFFFFF)) if ob & np.uint32(0x80000000) else ~ob
                    s.out_dist[i, j] = float(np.float32(1.0) - np.array([u], np.uint32).view(np.float32)[0])
                    s.out_keys[i, j] = cand[j][3]
                else:
                    s.out_dist[i, j] = float("inf")
                    s.out_keys[i, j] = -1

    # a searcher sized for a larger batch exchanges only the lists of the batch at hand and gives the same answer
    big = ShardedSearcher(FakeIndex(), K, B + 5, device=torch.device("cpu"), local_search=local_search, merge=merge)
    holder["s"] = big
    q = torch.from_numpy(queries)
    d_big, kk_big = big.search_device(q)
    assert big.local.shape == (2, B * K) and big.gathered.shape == (world, 2, B * K)
    d_big, kk_big = d_big.clone(), kk_big.clone()
    d_one, kk_one = big.search_device(q[:1])
    assert big.local.shape == (2, K) and torch.equal(kk_one[0], kk_big[0])
    assert torch.allclose(d_one[0], d_big[0], rtol=0, atol=1e-6)  # (numpy's gemv vs gemm: last-bit differences)
    s = ShardedSearcher(FakeIndex(), K, B, device=torch.device("cpu"), local_search=local_search, merge=merge)
    holder["s"] = s
    d, kk = s.search_device(q)
    d, kk = d.clone(), kk.clone()
    assert torch.equal(kk, kk_big) and torch.equal(d, d_big)
    # host path: only rank 0 owns the query block; the others receive it through the broadcast
    q_dev = torch.zeros_like(q)
    od, ok = torch.zeros((B, K)), torch.zeros((B, K), dtype=torch.int64)
    s.search_host(q if rank == 0 else None, q_dev, od, ok)
    if rank == 0:
        assert torch.equal(ok, kk) and torch.equal(od, d)
    np.save(os.path.join(out_dir, f"keys_{world}_{rank}.npy"), kk.numpy())
    np.save(os.path.join(out_dir, f"dist_{world}_{rank}.npy"), d.numpy())
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_search_equals_single_shard(world, tmp_path):
    from oracle import exact_scan as ox

    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    corpus = make_corpus(N, DIM, seed=5, dup_pairs=[(7, 4000), (8, 2600)])
    queries = make_queries(B, corpus, seed=6)
    queries[0], queries[1] = corpus[7], corpus[8]
    d1, r1 = ox.exact_topk(ox.prepare_queries(queries, "cosine"), ox.prepare_corpus(corpus, "cosine", "bf16"), K,
                           "cosine", "f32", prepared=True)
    for rank in range(world):
        kk = np.load(tmp_path / f"keys_{world}_{rank}.npy")
        dd = np.load(tmp_path / f"dist_{world}_{rank}.npy")
        np.testing.assert_array_equal(kk, r1)  # identical on every rank and equal to G = 1
        np.testing.assert_array_equal(dd, d1)
    assert r1[0, 0] == 7 and r1[0, 1] == 4000  # cross-shard tie -> lower global row first


def test_shard_bounds_cover_rows_exactly():
    from financial_rag_b200.sharded import shard_bounds

    for n, g, align in [(100_000_000, 8, 500_000), (10, 3, 1), (5003, 4, 100), (7, 8, 1), (0, 2, 1)]:
        spans = [shard_bounds(n, g, r, align) for r in range(g)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        for a, b in zip(spans, spans[1:]):
            assert a[1] == b[0] and a[0] <= a[1]
        assert all(lo % align == 0 for lo, _ in spans)
    assert shard_bounds(100_000_000, 8, 3, 500_000) == (37_500_000, 50_000_000)
