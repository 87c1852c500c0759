"""ShardGroup -- Python handle on one fr_group: ONE collection row-sharded over several B200s.

Same host surface as ``ShardIndex`` (upsert / delete / search / count / rows / get_rows / export_raw / import_raw /
lookup_rows), so ``B200Collection`` holds either without knowing which; a "row" here is a global row (insertion order
over the whole collection).  Everything below is ctypes over ``fr_group_*`` (include/fr_index.h); the scan, the
NCCL all-gather of the local top-k lists and the merge kernel are inside libfrb200.so.

Two ways to build one:
  * ``ShardGroup(devices=[0, 1, ..., 7])`` -- one process owns every shard (the reference's Flask server is one
    process: ``B200_CHILD_DEVICES=0,1,2,3,4,5,6,7`` or ``all`` makes ``get_child_vector_store`` hand out such a store);
  * ``ShardGroup.from_torch_distributed(device=local_rank)`` under ``torchrun`` -- one process per GPU, every rank
    makes the same calls (SPMD); torch.distributed only carries the 128-byte NCCL id at start-up.
"""
from __future__ import annotations

import ctypes
from typing import List, Optional, Sequence, Tuple

import numpy as np

from . import _lib
from ._lib import FR_XCHG_AUTO, FR_XCHG_COPY, FR_XCHG_NCCL, FR_XCHG_PEER, check
from .index import _DTYPES, _METRICS, _PATHS, ShardIndex, _stream_ptr, canonical_space

_EXCHANGE = {"auto": FR_XCHG_AUTO, "nccl": FR_XCHG_NCCL, "copy": FR_XCHG_COPY, "peer": FR_XCHG_PEER}
_EXCHANGE_NAME = {v: k for k, v in _EXCHANGE.items()}


def shard_of_row(row: int, world: int) -> Tuple[int, int]:
    """Cyclic placement: global row -> (shard, local row)."""
    return row % world, row // world


def rows_of_shard(shard: int, world: int, total_rows: int) -> int:
    """How many of the global rows [0, total_rows) live on ``shard``."""
    return (total_rows - shard + world - 1) // world if total_rows > shard else 0


def parse_devices(spec: Optional[str]) -> Optional[List[int]]:
    """``B200_CHILD_DEVICES``: "0,1,2,3" | "all" | "" (None = not set: one shard on B200_CHILD_DEVICE)."""
    if spec is None or not spec.strip():
        return None
    spec = spec.strip().lower()
    if spec == "all":
        import torch

        n = torch.cuda.device_count()
        if n < 1:
            raise RuntimeError("B200_CHILD_DEVICES=all but no CUDA device is visible (the backend has no CPU fallback)")
        return list(range(n))
    return [int(x) for x in spec.split(",") if x.strip()]


def exchange_nccl_id(device: Optional[int] = None, group=None) -> bytes:
    """Rank 0 of an initialised torch.distributed group creates an NCCL unique id (fr_nccl_unique_id) and every rank
    receives it: the only thing torch.distributed carries for a multi-process ShardGroup.  ``device``: CUDA ordinal
    when the process group's backend is NCCL (its broadcast wants device tensors); None for gloo."""
    import torch
    import torch.distributed as dist

    _lib.ensure_nccl()
    buf = ctypes.create_string_buffer(128)
    if dist.get_rank(group) == 0:
        check(_lib.load().fr_nccl_unique_id(buf, 128))
    t = torch.frombuffer(bytearray(buf.raw), dtype=torch.uint8).clone()
    if dist.get_backend(group) == "nccl":
        t = t.to(torch.device("cuda", device))
    dist.broadcast(t, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
    return bytes(t.cpu().numpy().tobytes())


class ShardGroup:
    def __init__(self, dim: int = 384, space: str = "cosine", dtype: str = "bf16", devices: Sequence[int] = (0,),
                 reserve_rows: int = 0, exchange: str = "auto", world_shards: Optional[int] = None,
                 first_shard: int = 0, nccl_id: Optional[bytes] = None):
        self._lib = _lib.load()
        self.dim = int(dim)
        self.space = canonical_space(space)
        if dtype not in _DTYPES:
            raise ValueError(f"dtype must be 'bf16' or 'f32', got {dtype!r}")
        if exchange not in _EXCHANGE:
            raise ValueError(f"exchange must be one of {sorted(_EXCHANGE)}, got {exchange!r}")
        self.dtype = dtype
        self.devices = [int(d) for d in devices]
        if not self.devices:
            raise ValueError("a group needs at least one device")
        self.world = int(world_shards) if world_shards else len(self.devices)
        self.first_shard = int(first_shard)
        if _EXCHANGE[exchange] == FR_XCHG_NCCL or (world_shards and world_shards != len(self.devices)):
            _lib.ensure_nccl()
        dev = (ctypes.c_int * len(self.devices))(*self.devices)
        per_shard = rows_of_shard(0, self.world, int(reserve_rows))
        idbuf = ctypes.create_string_buffer(bytes(nccl_id), 128) if nccl_id is not None else None
        h = ctypes.c_void_p()
        check(self._lib.fr_group_create(self.dim, _METRICS[self.space], _DTYPES[dtype], dev, len(self.devices),
                                        self.world, self.first_shard, idbuf, _EXCHANGE[exchange], per_shard,
                                        ctypes.byref(h)))
        self._h = h
        self._shards: List[ShardIndex] = []
        for j, d in enumerate(self.devices):
            sh = ctypes.c_void_p()
            check(self._lib.fr_group_shard(h, j, ctypes.byref(sh)))
            self._shards.append(ShardIndex._borrowed(sh, self.dim, self.space, self.dtype, d))
        self.device = self.devices[0]  # where host-API results are merged

    @classmethod
    def from_torch_distributed(cls, dim: int = 384, space: str = "cosine", dtype: str = "bf16", device: int = 0,
                               reserve_rows: int = 0, group=None) -> "ShardGroup":
        """One shard per rank of an initialised torch.distributed group (torchrun); rank r owns world shard r.
        The only thing torch.distributed moves is rank 0's NCCL unique id."""
        import torch.distributed as dist

        world, rank = dist.get_world_size(group), dist.get_rank(group)
        if world == 1:
            return cls(dim, space, dtype, [device], reserve_rows)
        return cls(dim, space, dtype, [device], reserve_rows, exchange="nccl", world_shards=world, first_shard=rank,
                   nccl_id=exchange_nccl_id(device, group))

    # -- lifecycle / introspection -------------------------------------------------------------------------
    def close(self) -> None:
        h, self._h = getattr(self, "_h", None), None
        if h:
            for s in self._shards:
                s._h = None
            self._lib.fr_group_destroy(h)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _handle(self):
        if not self._h:
            raise RuntimeError("ShardGroup is closed")
        return self._h

    def info(self, name: str) -> int:
        out = ctypes.c_int64()
        check(self._lib.fr_group_info(self._handle(), name.encode(), ctypes.byref(out)))
        return int(out.value)

    @property
    def exchange(self) -> str:
        return _EXCHANGE_NAME[self.info("exchange")]

    @property
    def n_local(self) -> int:
        return len(self.devices)

    def shard(self, j: int) -> ShardIndex:
        """Local shard ``j`` (world shard ``first_shard + j``) as a ShardIndex view: bulk device loads, options,
        statistics, scan profiling.  Call ``adopt_rows`` after loading rows through it."""
        self._handle()
        return self._shards[j]

    def adopt_rows(self, total_rows: int) -> None:
        check(self._lib.fr_group_adopt_rows(self._handle(), int(total_rows)))

    def reserve(self, total_rows: int) -> None:
        check(self._lib.fr_group_reserve(self._handle(), int(total_rows)))

    def set_option(self, name: str, value: int) -> None:
        check(self._lib.fr_group_set_option(self._handle(), name.encode(), int(value)))

    def set_path(self, path: str) -> None:
        self.set_option("path", _PATHS[path])

    def set_profile(self, on: bool) -> None:
        self.set_option("profile", 1 if on else 0)

    def profile_read(self):
        """Per local shard: (summed scan-kernel ms, scan launches, searches) since the last read."""
        return [s.profile_read() for s in self._shards]

    def stat(self, name: str) -> int:
        """Sum of a per-shard counter over the local shards (see ShardIndex.stat)."""
        return sum(s.stat(name) for s in self._shards)

    def count(self) -> int:
        return self.info("count")

    def rows(self) -> int:
        return self.info("rows")

    @property
    def row_bytes(self) -> int:
        return self.dim * (2 if self.dtype == "bf16" else 4)

    # -- host entry points (same contracts as ShardIndex) ------------------------------------------------------
    def upsert(self, vectors, keys) -> None:
        v = np.ascontiguousarray(vectors, dtype=np.float32)
        if v.ndim == 1:
            v = v[None, :]
        k = np.ascontiguousarray(keys, dtype=np.int64).reshape(-1)
        if v.ndim != 2 or v.shape[1] != self.dim:
            raise ValueError(f"expected vectors of shape (n, {self.dim}), got {v.shape}")
        if k.shape[0] != v.shape[0]:
            raise ValueError("one key per vector required")
        check(self._lib.fr_group_upsert(self._handle(), v.ctypes.data, k.ctypes.data, v.shape[0]))

    def delete(self, keys) -> int:
        k = np.ascontiguousarray(keys, dtype=np.int64).reshape(-1)
        out = ctypes.c_int64()
        check(self._lib.fr_group_delete(self._handle(), k.ctypes.data, k.shape[0], ctypes.byref(out)))
        return int(out.value)

    def search(self, queries, k: int) -> Tuple[np.ndarray, np.ndarray]:
        q = np.ascontiguousarray(queries, dtype=np.float32)
        if q.ndim == 1:
            q = q[None, :]
        if q.ndim != 2 or q.shape[1] != self.dim:
            raise ValueError(f"expected queries of shape (B, {self.dim}), got {q.shape}")
        b = q.shape[0]
        dist = np.empty((b, k), dtype=np.float32)
        keys = np.empty((b, k), dtype=np.int64)
        check(self._lib.fr_group_search(self._handle(), q.ctypes.data, b, int(k), dist.ctypes.data, keys.ctypes.data))
        return dist, keys

    def search_raw(self, q_ptr: Optional[int], b: int, k: int, dist_ptr: int, keys_ptr: int) -> None:
        """Same call on caller-owned host buffers.  ``q_ptr`` may be None on processes that do not own shard 0."""
        check(self._lib.fr_group_search(self._handle(), q_ptr, int(b), int(k), dist_ptr, keys_ptr))

    def get_rows(self, first_row: int, n: int) -> Tuple[np.ndarray, np.ndarray]:
        vecs = np.empty((n, self.dim), dtype=np.float32)
        keys = np.empty((n,), dtype=np.int64)
        check(self._lib.fr_group_get_rows(self._handle(), int(first_row), int(n), vecs.ctypes.data, keys.ctypes.data))
        return vecs, keys

    def export_raw(self, first_row: int, n: int) -> Tuple[np.ndarray, np.ndarray]:
        rows = np.empty((n, self.row_bytes), dtype=np.uint8)
        keys = np.empty((n,), dtype=np.int64)
        check(self._lib.fr_group_export_raw(self._handle(), int(first_row), int(n), rows.ctypes.data, keys.ctypes.data))
        return rows, keys

    def import_raw(self, rows, keys) -> None:
        k = np.ascontiguousarray(keys, dtype=np.int64).reshape(-1)
        r = np.asarray(rows)
        if r.dtype != np.uint8 or r.ndim != 2 or r.shape[1] != self.row_bytes or r.shape[0] != k.shape[0]:
            raise ValueError(f"expected uint8 rows of shape ({k.shape[0]}, {self.row_bytes}), got {r.dtype} {r.shape}")
        if not r.flags.c_contiguous:
            r = np.ascontiguousarray(r)
        check(self._lib.fr_group_import_raw(self._handle(), r.ctypes.data, k.ctypes.data, k.shape[0]))

    def lookup_rows(self, keys) -> np.ndarray:
        k = np.ascontiguousarray(keys, dtype=np.int64).reshape(-1)
        out = np.empty_like(k)
        check(self._lib.fr_group_lookup_rows(self._handle(), k.ctypes.data, k.shape[0], out.ctypes.data))
        return out

    # -- device entry point ------------------------------------------------------------------------------------
    def search_device(self, queries: Sequence, k: int, out_dist: Optional[Sequence] = None,
                      out_keys: Optional[Sequence] = None, streams: Optional[Sequence] = None, merge_on=None):
        """``queries[j]``: the [B, dim] fp32 block on the device of local shard j (the same block everywhere).
        Returns per-local-shard lists (dist, keys); entries are None where no merged result was asked for
        (``merge_on`` = local shard indices that want it, default: all)."""
        import torch

        n = self.n_local
        if len(queries) != n:
            raise ValueError(f"one query block per local shard required ({n}), got {len(queries)}")
        b = int(queries[0].shape[0])
        want = set(range(n)) if merge_on is None else set(merge_on)
        out_dist = list(out_dist) if out_dist is not None else [None] * n
        out_keys = list(out_keys) if out_keys is not None else [None] * n
        qp, dp, kp, sp = ((ctypes.c_void_p * n)() for _ in range(4))
        for j in range(n):
            self._shards[j]._check_dev(queries[j], torch.float32, f"queries[{j}]")
            if int(queries[j].shape[0]) != b:
                raise ValueError("every device must hold the same query block")
            if j in want:
                if out_dist[j] is None:
                    out_dist[j] = torch.empty((b, k), dtype=torch.float32, device=queries[j].device)
                if out_keys[j] is None:
                    out_keys[j] = torch.empty((b, k), dtype=torch.int64, device=queries[j].device)
            qp[j] = queries[j].data_ptr()
            dp[j] = out_dist[j].data_ptr() if out_dist[j] is not None else None
            kp[j] = out_keys[j].data_ptr() if out_keys[j] is not None else None
            sp[j] = _stream_ptr(streams[j] if streams is not None else None, self.devices[j])
        check(self._lib.fr_group_search_device(self._handle(), qp, b, int(k), dp, kp, sp))
        return out_dist, out_keys


__all__ = ["ShardGroup", "exchange_nccl_id", "shard_of_row", "rows_of_shard", "parse_devices"]
