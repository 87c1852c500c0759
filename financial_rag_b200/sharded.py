"""Row-sharded exact search: one process per GPU, NCCL all-gather of the local top-k, merge kernel.

SURVEY.md 8e: rank g of G holds the contiguous global rows [g*N/G, (g+1)*N/G) of the collection.
A search is
    1. every rank has the query block (rank 0's block is broadcast when it arrives on the host),
    2. each rank runs the scan + fused top-k on its shard  -> [B,k] packed keys + int64 keys,
    3. ONE all-gather of a [2, B, k] int64 buffer per rank (B*k*16 bytes; 160 B at B=1, k=10),
    4. every rank merges the G lists (ties -> lower shard, then lower local row == global
       insertion order), so the result is identical for any G.
The scan reads 768 B per row and exchanges 16 B per (query, result): the exchange is a latency
cost, not a bandwidth one, which is why it is a plain NCCL collective and not a fused kernel.

The two compute steps are injectable so the rank/partition/all-gather plumbing can be exercised
with the ``gloo`` backend on CPU (tests/test_sharded_gloo.py injects the CPU oracle there); the
defaults are the CUDA entry points and nothing else.
"""
from __future__ import annotations

from typing import Callable, Optional, Tuple


def shard_bounds(n_rows: int, world_size: int, rank: int, align: int = 1) -> Tuple[int, int]:
    """Contiguous row range of ``rank``; boundaries are multiples of ``align`` (except the end)."""
    units = (n_rows + align - 1) // align
    lo = (units * rank // world_size) * align
    hi = (units * (rank + 1) // world_size) * align
    return min(lo, n_rows), min(hi, n_rows)


class ShardedSearcher:
    def __init__(self, local_index, k: int, max_batch: int, *, space: str = "cosine", group=None,
                 device=None, local_search: Optional[Callable] = None, merge: Optional[Callable] = None):
        import torch
        import torch.distributed as dist

        self.dist = dist
        self.torch = torch
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.ix = local_index
        self.k = int(k)
        self.space = space
        self.max_batch = int(max_batch)
        if device is None:
            device = torch.device("cuda", local_index.device)
        self.device = torch.device(device)
        n = self.max_batch * self.k
        # [packed | keys] per rank, gathered into [G][2][B*k]; `local` / `gathered` are views sized for the batch of
        # the current call, so a batch of 1 exchanges 160 bytes per rank, not max_batch times that
        self._local_flat = torch.zeros((2 * n,), dtype=torch.int64, device=self.device)
        self._gathered_flat = torch.zeros((self.world * 2 * n,), dtype=torch.int64, device=self.device)
        self._views(self.max_batch)
        self.out_dist = torch.empty((self.max_batch, self.k), dtype=torch.float32, device=self.device)
        self.out_keys = torch.empty((self.max_batch, self.k), dtype=torch.int64, device=self.device)
        self._local_search = local_search or self._cuda_local_search
        self._merge = merge or self._cuda_merge

    def _views(self, b: int) -> None:
        m = b * self.k
        self.local = self._local_flat[: 2 * m].view(2, m)
        self.gathered = self._gathered_flat[: self.world * 2 * m].view(self.world, 2, m)

    # -- default (product) compute steps: CUDA through the C ABI -----------------------------------
    def _cuda_local_search(self, queries, b: int) -> None:
        self.ix.search_partial_device(queries, self.k, self.local[0], self.local[1])

    def _cuda_merge(self, b: int) -> None:
        from .index import merge_shards_device

        merge_shards_device(self.device.index, self.space, self.gathered[:, 0], self.gathered[:, 1], 2 * b * self.k,
                            self.world, b, self.k, self.out_dist, self.out_keys)

    # -- the search -----------------------------------------------------------------------------------
    def search_device(self, queries):
        """``queries``: [B, dim] fp32 on this rank's device, identical on every rank.
        Returns (dist [B,k], keys [B,k]) device tensors, identical on every rank."""
        b = int(queries.shape[0])
        if b > self.max_batch:
            raise ValueError(f"batch {b} exceeds max_batch {self.max_batch}")
        self._views(b)
        self._local_search(queries, b)
        if self.world > 1:
            self.dist.all_gather_into_tensor(self.gathered.view(-1), self.local.view(-1), group=self.group)
        else:
            self.gathered[0].copy_(self.local)
        self._merge(b)
        return self.out_dist[:b], self.out_keys[:b]

    def search_host(self, q_host_pinned, q_dev, out_dist_host, out_keys_host):
        """End-to-end form: rank 0 holds the query block in pinned host memory; it is copied to the
        device, broadcast over NCCL, searched, and the result lands in rank 0's pinned buffers."""
        b = int(q_dev.shape[0])
        if self.rank == 0:
            q_dev.copy_(q_host_pinned, non_blocking=True)
        if self.world > 1:
            self.dist.broadcast(q_dev, src=0, group=self.group)
        d, kk = self.search_device(q_dev)
        if self.rank == 0:
            out_dist_host[:b].copy_(d, non_blocking=True)
            out_keys_host[:b].copy_(kk, non_blocking=True)
        if self.device.type == "cuda":
            self.torch.cuda.current_stream(self.device).synchronize()
        return out_dist_host, out_keys_host
