"""EnsembleSearcher -- the dual-encoder ensemble of the reference, device-resident from the query block to the fused
top-k (BASELINE.json cfg3; north star: "followed by on-GPU reciprocal-rank/score fusion across the per-encoder
collections").

The reference (parent_child/retriever.py:80-107; rag_backend.py:653-731) embeds the query once per encoder, asks each
``children_<slug(encoder)>`` collection for its top ``k_each`` and fuses the ranked lists in Python.  Here the L
collections -- each one shard (``ShardIndex``) or a row-sharded group (``ShardGroup``) -- are scanned back to back on
the device, their merged key (and distance) lists land in one ``[L, B, k_each]`` buffer, and ONE fusion kernel turns
them into the fused top-``k_out``: K5 ``rrf_fuse_kernel`` (``fusion="rrf"``, the reference's only live mode,
rag_backend.py:589) or ``score_fuse_kernel`` (``fusion="avg"``: per-list min-max normalised scores, mean over lists,
rag_backend.py:732-754).  Nothing but the query blocks goes in and nothing but ``[B, k_out]`` comes out.
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import numpy as np

from .group import ShardGroup
from .index import rrf_fuse_device, score_fuse_device


class EnsembleSearcher:
    def __init__(self, indexes: Sequence, k_each: int = 50, k_rrf: int = 60, k_out: int = 10, fusion: str = "rrf"):
        if not indexes:
            raise ValueError("an ensemble needs at least one collection")
        if fusion not in ("rrf", "avg"):
            raise ValueError(f"fusion must be 'rrf' or 'avg', got {fusion!r}")
        self.indexes = list(indexes)
        self.k_each, self.k_rrf, self.k_out, self.fusion = int(k_each), int(k_rrf), int(k_out), fusion
        # the fused result is produced where every collection's merged lists arrive: the first local device
        self.device = self.indexes[0].device
        for ix in self.indexes:
            if ix.device != self.device:
                raise ValueError("the collections of an ensemble must merge on the same device")
        self._bufs = {}

    def _buffers(self, b: int):
        import torch

        buf = self._bufs.get(b)
        if buf is None:
            dev = torch.device("cuda", self.device)
            n = len(self.indexes)
            buf = (torch.empty((n, b, self.k_each), dtype=torch.int64, device=dev),
                   torch.empty((n, b, self.k_each), dtype=torch.float32, device=dev))
            if len(self._bufs) > 8:
                self._bufs.clear()
            self._bufs[b] = buf
        return buf

    def search_device(self, queries: Sequence, streams: Optional[Sequence] = None):
        """``queries[l]``: the query block of collection l as ITS encoder embedded it -- a CUDA tensor [B, dim] for a
        ``ShardIndex``, a list of them (one per local device) for a ``ShardGroup``.
        Returns CUDA tensors (fused score fp64 [B, k_out], keys int64 [B, k_out]; -1 pads) on ``self.device``."""
        if len(queries) != len(self.indexes):
            raise ValueError("one query block per collection required")
        first = queries[0][0] if isinstance(queries[0], (list, tuple)) else queries[0]
        b = int(first.shape[0])
        keys, dist = self._buffers(b)
        for l, (ix, q) in enumerate(zip(self.indexes, queries)):
            if isinstance(ix, ShardGroup):
                n = ix.n_local
                ix.search_device(list(q), self.k_each, [dist[l]] + [None] * (n - 1), [keys[l]] + [None] * (n - 1),
                                 streams=streams, merge_on=[0])
            else:
                ix.search_device(q, self.k_each, dist[l], keys[l], stream=streams[0] if streams else None)
        st = streams[0] if streams else None
        if self.fusion == "rrf":
            return rrf_fuse_device(keys, self.k_rrf, self.k_out, stream=st)
        return score_fuse_device(dist, keys, self.k_out, stream=st)

    def search(self, queries: Sequence) -> Tuple[np.ndarray, np.ndarray]:
        """Host buffers in (one [B, dim] fp32 array or CPU tensor per collection -- pinned ones are copied without a
        staging pass), host buffers out: (fused score fp64 [B, k_out], keys int64 [B, k_out])."""
        import torch

        blocks: List = []
        for ix, q in zip(self.indexes, queries):
            t = q if isinstance(q, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(q, dtype=np.float32))
            if isinstance(ix, ShardGroup):
                blocks.append([t.to(torch.device("cuda", d), non_blocking=True) for d in ix.devices])
            else:
                blocks.append(t.to(torch.device("cuda", ix.device), non_blocking=True))
        sc, keys = self.search_device(blocks)
        return sc.cpu().numpy(), keys.cpu().numpy()


__all__ = ["EnsembleSearcher"]
