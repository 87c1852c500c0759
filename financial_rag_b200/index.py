"""ShardIndex -- Python handle on one fr_index (one collection shard resident on one B200).

Thin host code over the C ABI (include/fr_index.h); all arithmetic happens in the CUDA kernels.
numpy arrays are used for the host entry points, torch tensors only as owners of device memory
for the ``*_device`` entry points.
"""
from __future__ import annotations

import ctypes
from typing import Optional, Tuple

import numpy as np

from . import _lib
from ._lib import FR_BF16, FR_COSINE, FR_F32, FR_IP, FR_L2, FR_MAX_K, check

_METRICS = {"cosine": FR_COSINE, "l2": FR_L2, "ip": FR_IP}
_DTYPES = {"bf16": FR_BF16, "f32": FR_F32}
_PATHS = {"auto": 0, "stream": 1, "mma": 2}


def canonical_space(space: Optional[str]) -> str:
    """Metric vocabulary of the reference (parent_child/pgvector_child_store.py:7-26);
    unknown names fall back to cosine exactly as ``_get_distance_ops`` does."""
    d = (space or "cosine").lower()
    if d in ("cos", "cosine"):
        return "cosine"
    if d in ("l2", "euclidean"):
        return "l2"
    if d in ("ip", "inner", "inner_product"):
        return "ip"
    return "cosine"


def _stream_ptr(stream, device: Optional[int] = None) -> ctypes.c_void_p:
    """``stream`` = a torch stream, a raw cudaStream_t, or None = torch's current stream ON ``device`` (the index's
    device, which need not be torch's current device)."""
    if stream is None:
        import torch

        stream = torch.cuda.current_stream(device)
    return ctypes.c_void_p(int(getattr(stream, "cuda_stream", stream)))


class ShardIndex:
    def __init__(self, dim: int = 384, space: str = "cosine", dtype: str = "bf16", device: int = 0,
                 reserve_rows: int = 0):
        self._lib = _lib.load()
        self.dim = int(dim)
        self.space = canonical_space(space)
        if dtype not in _DTYPES:
            raise ValueError(f"dtype must be 'bf16' or 'f32', got {dtype!r}")
        self.dtype = dtype
        self.device = int(device)
        h = ctypes.c_void_p()
        check(self._lib.fr_index_create(self.dim, _METRICS[self.space], _DTYPES[dtype], self.device,
                                        int(reserve_rows), ctypes.byref(h)))
        self._h = h
        self._owned = True

    @classmethod
    def _borrowed(cls, handle, dim: int, space: str, dtype: str, device: int) -> "ShardIndex":
        """View of a shard some fr_group owns (ShardGroup.shard): same calls, never destroyed from here."""
        self = cls.__new__(cls)
        self._lib = _lib.load()
        self.dim, self.space, self.dtype, self.device = int(dim), space, dtype, int(device)
        self._h = handle
        self._owned = False
        return self

    # -- lifecycle ---------------------------------------------------------------------------
    def close(self) -> None:
        h, self._h = getattr(self, "_h", None), None
        if h and getattr(self, "_owned", True):
            self._lib.fr_index_destroy(h)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _handle(self):
        if not self._h:
            raise RuntimeError("ShardIndex is closed")
        return self._h

    def reserve(self, rows: int) -> None:
        check(self._lib.fr_index_reserve(self._handle(), int(rows)))

    def set_path(self, path: str) -> None:
        """Pin the kernel regime: 'auto' | 'stream' (K1) | 'mma' (K2).  Tests and bench use it."""
        check(self._lib.fr_index_set_option(self._handle(), b"path", _PATHS[path]))

    def set_option(self, name: str, value: int) -> None:
        """Tuning knobs of the C ABI: 'mma_min_batch', 'mma_small_max', 'mma_co_groups', 'mma_split' (-1 auto, 0, 1),
        'mma_split_max', 'path', 'profile'."""
        check(self._lib.fr_index_set_option(self._handle(), name.encode(), int(value)))

    def stat(self, name: str) -> int:
        """'searches' | 'queries' | 'mma_queries' | 'mma_uncertified_queries' | 'mma_rescanned_queries' since creation."""
        out = ctypes.c_int64()
        check(self._lib.fr_index_get_stat(self._handle(), name.encode(), ctypes.byref(out)))
        return int(out.value)

    def set_profile(self, on: bool) -> None:
        """Bracket every scan launch with CUDA events (bench.py's roofline line)."""
        check(self._lib.fr_index_set_option(self._handle(), b"profile", 1 if on else 0))

    def profile_read(self):
        """(summed scan-kernel ms, scan launches, searches) since the last read; resets."""
        ms, nl, ns = ctypes.c_double(), ctypes.c_int64(), ctypes.c_int64()
        check(self._lib.fr_index_profile_read(self._handle(), ctypes.byref(ms), ctypes.byref(nl), ctypes.byref(ns)))
        return float(ms.value), int(nl.value), int(ns.value)

    def count(self) -> int:
        out = ctypes.c_int64()
        check(self._lib.fr_index_count(self._handle(), ctypes.byref(out)))
        return int(out.value)

    def rows(self) -> int:
        out = ctypes.c_int64()
        check(self._lib.fr_index_rows(self._handle(), ctypes.byref(out)))
        return int(out.value)

    # -- host entry points -----------------------------------------------------------------
    def upsert(self, vectors, keys) -> None:
        v = np.ascontiguousarray(vectors, dtype=np.float32)
        if v.ndim == 1:
            v = v[None, :]
        k = np.ascontiguousarray(keys, dtype=np.int64).reshape(-1)
        if v.ndim != 2 or v.shape[1] != self.dim:
            raise ValueError(f"expected vectors of shape (n, {self.dim}), got {v.shape}")
        if k.shape[0] != v.shape[0]:
            raise ValueError("one key per vector required")
        check(self._lib.fr_index_upsert(self._handle(), v.ctypes.data, k.ctypes.data, v.shape[0]))

    def delete(self, keys) -> int:
        k = np.ascontiguousarray(keys, dtype=np.int64).reshape(-1)
        out = ctypes.c_int64()
        check(self._lib.fr_index_delete(self._handle(), k.ctypes.data, k.shape[0], ctypes.byref(out)))
        return int(out.value)

    def search(self, queries, k: int) -> Tuple[np.ndarray, np.ndarray]:
        """Host buffers in, host buffers out (H2D/D2H inside): (dist [B,k] fp32, keys [B,k] int64)."""
        q = np.ascontiguousarray(queries, dtype=np.float32)
        if q.ndim == 1:
            q = q[None, :]
        if q.ndim != 2 or q.shape[1] != self.dim:
            raise ValueError(f"expected queries of shape (B, {self.dim}), got {q.shape}")
        b = q.shape[0]
        dist = np.empty((b, k), dtype=np.float32)
        keys = np.empty((b, k), dtype=np.int64)
        check(self._lib.fr_index_search(self._handle(), q.ctypes.data, b, int(k), dist.ctypes.data,
                                        keys.ctypes.data))
        return dist, keys

    def search_raw(self, q_ptr: int, b: int, k: int, dist_ptr: int, keys_ptr: int) -> None:
        """Same call on caller-owned host buffers (e.g. pinned torch tensors) -- no allocation."""
        check(self._lib.fr_index_search(self._handle(), q_ptr, int(b), int(k), dist_ptr, keys_ptr))

    def get_rows(self, first_row: int, n: int) -> Tuple[np.ndarray, np.ndarray]:
        vecs = np.empty((n, self.dim), dtype=np.float32)
        keys = np.empty((n,), dtype=np.int64)
        check(self._lib.fr_index_get_rows(self._handle(), int(first_row), int(n), vecs.ctypes.data,
                                          keys.ctypes.data))
        return vecs, keys

    # -- persistence: rows exactly as stored in HBM -------------------------------
    @property
    def row_bytes(self) -> int:
        return self.dim * (2 if self.dtype == "bf16" else 4)

    def export_raw(self, first_row: int, n: int) -> Tuple[np.ndarray, np.ndarray]:
        """(rows [n, row_bytes] uint8 -- the stored bf16/fp32 bits -- , keys [n] int64; INT64_MIN = deleted)."""
        rows = np.empty((n, self.row_bytes), dtype=np.uint8)
        keys = np.empty((n,), dtype=np.int64)
        check(self._lib.fr_index_export_raw(self._handle(), int(first_row), int(n), rows.ctypes.data, keys.ctypes.data))
        return rows, keys

    def import_raw(self, rows, keys) -> None:
        """Append rows verbatim (already prepared storage bits, e.g. a memory-mapped shard file)."""
        k = np.ascontiguousarray(keys, dtype=np.int64).reshape(-1)
        r = np.asarray(rows)
        if r.dtype != np.uint8 or r.ndim != 2 or r.shape[1] != self.row_bytes or r.shape[0] != k.shape[0]:
            raise ValueError(f"expected uint8 rows of shape ({k.shape[0]}, {self.row_bytes}), got {r.dtype} {r.shape}")
        if not r.flags.c_contiguous:
            r = np.ascontiguousarray(r)
        check(self._lib.fr_index_import_raw(self._handle(), r.ctypes.data, k.ctypes.data, k.shape[0]))

    def lookup_rows(self, keys) -> np.ndarray:
        """Physical row of each key (-1 = absent)."""
        k = np.ascontiguousarray(keys, dtype=np.int64).reshape(-1)
        out = np.empty_like(k)
        check(self._lib.fr_index_lookup_rows(self._handle(), k.ctypes.data, k.shape[0], out.ctypes.data))
        return out

    # -- device entry points (torch tensors own the memory) --------------------------------------
    def _check_dev(self, t, dtype, what):
        import torch

        if not (isinstance(t, torch.Tensor) and t.is_cuda and t.device.index == self.device):
            raise ValueError(f"{what} must be a CUDA tensor on device {self.device}")
        if t.dtype != dtype or not t.is_contiguous():
            raise ValueError(f"{what} must be contiguous {dtype}")

    def append_device(self, vectors, keys=None, first_key: int = 0, stream=None) -> None:
        import torch

        self._check_dev(vectors, torch.float32, "vectors")
        if vectors.ndim != 2 or vectors.shape[1] != self.dim:
            raise ValueError(f"expected vectors of shape (n, {self.dim})")
        kp = None
        if keys is not None:
            self._check_dev(keys, torch.int64, "keys")
            kp = keys.data_ptr()
        check(self._lib.fr_index_append_device(self._handle(), vectors.data_ptr(), kp, int(first_key),
                                               vectors.shape[0], _stream_ptr(stream, self.device)))

    def search_device(self, queries, k: int, out_dist=None, out_keys=None, stream=None):
        import torch

        self._check_dev(queries, torch.float32, "queries")
        b = queries.shape[0]
        if out_dist is None:
            out_dist = torch.empty((b, k), dtype=torch.float32, device=queries.device)
        if out_keys is None:
            out_keys = torch.empty((b, k), dtype=torch.int64, device=queries.device)
        check(self._lib.fr_index_search_device(self._handle(), queries.data_ptr(), b, int(k),
                                               out_dist.data_ptr(), out_keys.data_ptr(), _stream_ptr(stream, self.device)))
        return out_dist, out_keys

    def search_partial_device(self, queries, k: int, out_packed, out_keys, stream=None) -> None:
        """Local top-k in mergeable form (step 1 of the row-sharded search).
        out_packed / out_keys: int64 tensors of B*k elements (packed keys are raw uint64 bits)."""
        import torch

        self._check_dev(queries, torch.float32, "queries")
        check(self._lib.fr_index_search_partial_device(self._handle(), queries.data_ptr(), queries.shape[0],
                                                       int(k), out_packed.data_ptr(), out_keys.data_ptr(),
                                                       _stream_ptr(stream, self.device)))


def merge_shards_device(device: int, space: str, packed, keys, shard_stride: int, g: int, b: int, k: int,
                        out_dist, out_keys, stream=None) -> None:
    """Step 2 of the row-sharded search: merge G gathered lists (K3 in SHARDS mode)."""
    lib = _lib.load()
    check(lib.fr_merge_shards_device(int(device), _METRICS[canonical_space(space)], packed.data_ptr(),
                                     keys.data_ptr(), int(shard_stride), int(g), int(b), int(k),
                                     out_dist.data_ptr(), out_keys.data_ptr(), _stream_ptr(stream, int(device))))


def rrf_fuse_host(keys: np.ndarray, k_rrf: int = 60, k_out: int = 10, device: int = 0):
    """keys: [L, B, kp] int64 (-1 = empty).  Returns (score [B,k_out] fp64, keys [B,k_out])."""
    lib = _lib.load()
    a = np.ascontiguousarray(keys, dtype=np.int64)
    if a.ndim != 3:
        raise ValueError("keys must be [L, B, kp]")
    l, b, kp = a.shape
    sc = np.zeros((b, k_out), dtype=np.float64)
    ok = np.full((b, k_out), -1, dtype=np.int64)
    check(lib.fr_rrf_fuse(int(device), a.ctypes.data, l, b, kp, int(k_rrf), int(k_out), sc.ctypes.data,
                          ok.ctypes.data))
    return sc, ok


def maxsim_aggregate_host(dist: np.ndarray, keys: np.ndarray, group_shift: int, k_out: int = 24, device: int = 0):
    """dist/keys: [B, T, kp] per-token hit lists of B queries.  Returns (score [B,k_out] fp64, group [B,k_out])."""
    lib = _lib.load()
    d = np.ascontiguousarray(dist, dtype=np.float32)
    k = np.ascontiguousarray(keys, dtype=np.int64)
    if d.ndim != 3 or d.shape != k.shape:
        raise ValueError("dist and keys must both be [B, T, kp]")
    b, t, kp = d.shape
    sc = np.zeros((b, k_out), dtype=np.float64)
    og = np.full((b, k_out), -1, dtype=np.int64)
    check(lib.fr_maxsim_aggregate(int(device), d.ctypes.data, k.ctypes.data, b, t, kp, int(group_shift), int(k_out),
                                  sc.ctypes.data, og.ctypes.data))
    return sc, og


def maxsim_aggregate_device(dist, keys, group_shift: int, k_out: int = 24, stream=None):
    """CUDA tensors [B, T, kp] (fp32 dist, int64 keys) -> CUDA tensors (score fp64 [B,k_out], group [B,k_out])."""
    import torch

    lib = _lib.load()
    if not (dist.is_cuda and keys.is_cuda and dist.dtype == torch.float32 and keys.dtype == torch.int64
            and dist.is_contiguous() and keys.is_contiguous() and dist.ndim == 3 and dist.shape == keys.shape):
        raise ValueError("dist (fp32) and keys (int64) must be contiguous CUDA tensors [B, T, kp]")
    b, t, kp = dist.shape
    sc = torch.empty((b, k_out), dtype=torch.float64, device=dist.device)
    og = torch.empty((b, k_out), dtype=torch.int64, device=dist.device)
    check(lib.fr_maxsim_aggregate_device(dist.device.index, dist.data_ptr(), keys.data_ptr(), b, t, kp, int(group_shift),
                                         int(k_out), sc.data_ptr(), og.data_ptr(), _stream_ptr(stream, dist.device.index)))
    return sc, og


def rrf_fuse_device(keys, k_rrf: int = 60, k_out: int = 10, stream=None):
    """keys: CUDA int64 tensor [L, B, kp].  Returns CUDA tensors (score fp64 [B,k_out], keys)."""
    import torch

    lib = _lib.load()
    if not (keys.is_cuda and keys.dtype == torch.int64 and keys.is_contiguous() and keys.ndim == 3):
        raise ValueError("keys must be a contiguous CUDA int64 tensor [L, B, kp]")
    l, b, kp = keys.shape
    sc = torch.empty((b, k_out), dtype=torch.float64, device=keys.device)
    ok = torch.empty((b, k_out), dtype=torch.int64, device=keys.device)
    check(lib.fr_rrf_fuse_device(keys.device.index, keys.data_ptr(), l, b, kp, int(k_rrf), int(k_out),
                                 sc.data_ptr(), ok.data_ptr(), _stream_ptr(stream, keys.device.index)))
    return sc, ok


def score_fuse_host(dist: np.ndarray, keys: np.ndarray, k_out: int = 10, device: int = 0):
    """``avg`` fusion (rag_backend.py:732-754): dist/keys [L, B, kp] -> (score [B,k_out] fp64, keys [B,k_out])."""
    lib = _lib.load()
    d = np.ascontiguousarray(dist, dtype=np.float32)
    k = np.ascontiguousarray(keys, dtype=np.int64)
    if d.ndim != 3 or d.shape != k.shape:
        raise ValueError("dist and keys must both be [L, B, kp]")
    l, b, kp = d.shape
    sc = np.zeros((b, k_out), dtype=np.float64)
    ok = np.full((b, k_out), -1, dtype=np.int64)
    check(lib.fr_score_fuse(int(device), d.ctypes.data, k.ctypes.data, l, b, kp, int(k_out), sc.ctypes.data, ok.ctypes.data))
    return sc, ok


def score_fuse_device(dist, keys, k_out: int = 10, stream=None):
    """CUDA tensors [L, B, kp] (fp32 dist, int64 keys) -> CUDA tensors (score fp64 [B,k_out], keys [B,k_out])."""
    import torch

    lib = _lib.load()
    if not (dist.is_cuda and keys.is_cuda and dist.dtype == torch.float32 and keys.dtype == torch.int64
            and dist.is_contiguous() and keys.is_contiguous() and dist.ndim == 3 and dist.shape == keys.shape):
        raise ValueError("dist (fp32) and keys (int64) must be contiguous CUDA tensors [L, B, kp]")
    l, b, kp = dist.shape
    sc = torch.empty((b, k_out), dtype=torch.float64, device=dist.device)
    ok = torch.empty((b, k_out), dtype=torch.int64, device=dist.device)
    check(lib.fr_score_fuse_device(dist.device.index, dist.data_ptr(), keys.data_ptr(), l, b, kp, int(k_out),
                                   sc.data_ptr(), ok.data_ptr(), _stream_ptr(stream, dist.device.index)))
    return sc, ok


__all__ = ["ShardIndex", "score_fuse_host", "score_fuse_device", "canonical_space", "merge_shards_device", "rrf_fuse_host", "rrf_fuse_device",
           "maxsim_aggregate_host", "maxsim_aggregate_device", "FR_MAX_K"]
