"""Run under torchrun by tests/test_gpu_group.py::test_multi_process_group_under_torchrun (one process per GPU).

Every rank builds the same synthetic collection through the SPMD group (each keeps its own rows), searches it through
the host and the device entry points and compares with a one-GPU index holding all rows on its own device.
Exits non-zero on any mismatch on any rank.
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main() -> int:
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    import financial_rag_b200 as frb
    from helpers import make_corpus, make_queries

    n = 50000
    corpus = make_corpus(n, seed=1, dup_pairs=[(0, n - 1), (3, 40000)])
    keys = np.arange(n, dtype=np.int64) + 77
    grp = frb.ShardGroup.from_torch_distributed(dim=384, dtype="bf16", device=local)
    assert grp.world == world and grp.first_shard == rank and grp.exchange == "nccl"
    one = frb.ShardIndex(dim=384, dtype="bf16", device=local)
    for lo in range(0, n, 9001):  # SPMD: the same upserts on every rank
        grp.upsert(corpus[lo:lo + 9001], keys[lo:lo + 9001])
        one.upsert(corpus[lo:lo + 9001], keys[lo:lo + 9001])
    assert grp.delete(keys[[5, 6, 7]]) == 3 and one.delete(keys[[5, 6, 7]]) == 3
    assert grp.count() == n - 3 and grp.shard(0).rows() == len(range(rank, n, world))
    ok = True
    for B, k in ((1, 10), (50, 10), (300, 10), (130, 100)):
        q = make_queries(B, corpus, seed=B)
        q[0] = corpus[0]
        wd, wk = one.search(q, k)
        # host form: only the owner of shard 0 feeds the queries
        gd = np.empty((B, k), np.float32)
        gk = np.empty((B, k), np.int64)
        qq = np.ascontiguousarray(q)
        grp.search_raw(qq.ctypes.data if rank == 0 else None, B, k, gd.ctypes.data, gk.ctypes.data)
        fin = np.isfinite(wd)
        same = bool((gk == wk).all() and np.abs(gd[fin] - wd[fin]).max() <= 2e-6)
        same &= gk[0, 0] == 77 and gk[0, 1] == 77 + n - 1
        # device form
        dd, dk = grp.search_device([torch.from_numpy(q).to(dev)], k)
        torch.cuda.synchronize()
        same &= bool((dk[0].cpu().numpy() == wk).all())
        if not same:
            print(f"[rank {rank}] MISMATCH at B={B} k={k}", flush=True)
        ok &= same
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    grp.close()
    one.close()
    dist.destroy_process_group()
    if rank == 0 and int(flag.item()) == 1:
        print(f"SPMD-OK world={world}", flush=True)
    return 0 if int(flag.item()) == 1 else 1


if __name__ == "__main__":
    sys.exit(main())
