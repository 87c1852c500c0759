"""World-size-2 / 3 runs, on CPU with the gloo backend, of the host-side plumbing of the row-sharded path:

  * the cyclic placement arithmetic (``shard_of_row`` / ``rows_of_shard``) and the merge order it implies -- every
    rank holds the rows r with r % W == rank, computes its local top-k with the CPU oracle (in the TEST only), the
    lists are all-gathered and merged by (score, local_row * W + shard) exactly as merge_topk.cu's cyclic mode does:
    the result must equal the one-shard oracle, ties in global insertion order;
  * ``exchange_nccl_id``: rank 0's NCCL unique id reaches every rank (the one thing torch.distributed carries for a
    multi-process fr_group), and without a CUDA device the group then refuses loudly (no CPU fallback);
  * bench.py's checker (``fp32_reference`` + ``north_star_gate``): chunks dealt round-robin to the ranks, per-rank
    top-k all-gathered and merged -- against a plain numpy brute force.
"""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from helpers import make_corpus, make_queries  # noqa: E402

N, B, K, DIM = 4001, 5, 10, 384


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import exact_scan as ox

    from financial_rag_b200 import _lib
    from financial_rag_b200.group import ShardGroup, exchange_nccl_id, rows_of_shard, shard_of_row

    # ---- cyclic placement + merge order --------------------------------------------------------------------------
    corpus = make_corpus(N, DIM, seed=5, dup_pairs=[(7, 4000), (8, 2601), (9, 9 + world)])
    queries = make_queries(B, corpus, seed=6)
    queries[0], queries[1], queries[2] = corpus[7], corpus[8], corpus[9]
    mine = np.arange(rank, N, world)
    assert len(mine) == rows_of_shard(rank, world, N)
    assert all(shard_of_row(int(r), world) == (rank, i) for i, r in enumerate(mine[:50]))
    shard = ox.prepare_corpus(corpus[mine], "cosine", "bf16")
    qp = ox.prepare_queries(queries, "cosine")
    d, r = ox.exact_topk(qp, shard, K, "cosine", "f32", prepared=True)       # local rows, local order
    local = torch.from_numpy(np.stack([d.astype(np.float64), r.astype(np.float64)]))
    gathered = [torch.empty_like(local) for _ in range(world)]
    dist.all_gather(gathered, local)
    got_rows = np.empty((B, K), np.int64)
    got_dist = np.empty((B, K), np.float32)
    for i in range(B):
        cand = []
        for s, g in enumerate(gathered):
            for dd, rr in zip(g[0, i].tolist(), g[1, i].tolist()):
                if rr >= 0:
                    cand.append((np.float32(dd), int(rr) * world + s))         # (distance, local_row * W + shard)
        cand.sort()
        got_dist[i] = [c[0] for c in cand[:K]]
        got_rows[i] = [c[1] for c in cand[:K]]                                 # local_row * W + shard IS the global row
    want_d, want_r = ox.exact_topk(qp, ox.prepare_corpus(corpus, "cosine", "bf16"), K, "cosine", "f32", prepared=True)
    assert (got_rows == want_r).all(), (got_rows[:3], want_r[:3])
    assert np.abs(got_dist - want_d).max() <= 1e-6
    assert got_rows[0, :2].tolist() == [7, 4000] and got_rows[2, :2].tolist() == [9, 9 + world]  # ties: insertion order

    # ---- the NCCL id reaches every rank; no GPU here, so the group must refuse loudly ----------------------------
    uid = exchange_nccl_id(None)
    assert len(uid) == 128 and any(uid)
    ids = [None] * world
    dist.all_gather_object(ids, uid)
    assert all(x == ids[0] for x in ids)
    if not torch.cuda.is_available():
        with pytest.raises(_lib.FrError) as ei:
            ShardGroup.from_torch_distributed(dim=DIM, device=0)
        assert ei.value.code == -4 and "no CPU fallback" in str(ei.value)

    # ---- bench.py's distributed checker ---------------------------------------------------------------------------
    import bench

    old_chunk = bench.CHUNK_ROWS
    bench.CHUNK_ROWS = 1000
    try:
        n = 4500
        cpu = torch.device("cpu")
        q_raw = torch.randn((6, DIM), generator=torch.Generator().manual_seed(1))
        full = torch.cat([bench.gen_chunk(torch, cpu, 77, c, min(1000, n - c * 1000)) for c in range(5)])
        fulln = full / (full.norm(dim=1, keepdim=True) + 1e-30)
        qn = q_raw / (q_raw.norm(dim=1, keepdim=True) + 1e-30)
        s = (qn @ fulln.T).numpy()
        order = np.argsort(-s, axis=1, kind="stable")[:, :K]
        got = order.copy()
        got[0, 3] = 4321                                    # the "GPU" returned a wrong row here
        ref = bench.fp32_reference(torch, dist, cpu, rank, world, n, 77, q_raw, K, got)
        ref_k, ref_s, _, got_s, _ = ref
        assert (ref_k == order).all()
        np.testing.assert_allclose(ref_s, np.take_along_axis(s, order, 1), atol=1e-6)
        np.testing.assert_allclose(got_s, np.take_along_axis(s, got, 1), atol=1e-6)
        gate = bench.north_star_gate(got, (1.0 - np.take_along_axis(s, got, 1)).astype(np.float32), ref, "f32", n)
        assert gate["id_mismatches"] == 1 and gate["id_mismatches_outside_1e-3_ties"] == 1 and not gate["pass"]
        ok = bench.north_star_gate(order, (1.0 - np.take_along_axis(s, order, 1)).astype(np.float32),
                                   bench.fp32_reference(torch, dist, cpu, rank, world, n, 77, q_raw, K, order), "f32", n)
        assert ok["pass"] and ok["pass_literal"] and ok["id_mismatches"] == 0
    finally:
        bench.CHUNK_ROWS = old_chunk
    with open(os.path.join(out_dir, f"ok_{world}_{rank}"), "w") as f:
        f.write("ok")
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_cyclic_sharding_plumbing_under_gloo(world, tmp_path):
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    assert all(os.path.exists(os.path.join(str(tmp_path), f"ok_{world}_{r}")) for r in range(world))


def test_cyclic_placement_arithmetic():
    from financial_rag_b200.group import parse_devices, rows_of_shard, shard_of_row

    for world in (1, 2, 3, 8):
        for total in (0, 1, 7, 8, 9, 1000, 100_000_001):
            assert sum(rows_of_shard(s, world, total) for s in range(world)) == total
            counts = [rows_of_shard(s, world, total) for s in range(world)]
            assert max(counts) - min(counts) <= 1                               # balanced to within one row, always
            if total:
                s, l = shard_of_row(total - 1, world)
                assert l == rows_of_shard(s, world, total) - 1                  # the last row is its shard's last row
    assert parse_devices(None) is None and parse_devices("") is None
    assert parse_devices("0,1, 3") == [0, 1, 3] and parse_devices("2") == [2]
