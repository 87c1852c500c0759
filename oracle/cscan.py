"""ctypes binding of oracle/exact_scan.c (CPU ORACLE -- test infrastructure, not product code)."""
from __future__ import annotations

import ctypes
import hashlib
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "liboracle_scan.so")
_SPACE = {"cosine": 0, "l2": 1, "ip": 2}
_lib = None


def _cpu_stamp() -> str:
    """-march=native code must be rebuilt when the .so travels to a box with another CPU."""
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.startswith("flags"):
                    return hashlib.sha1(line.encode()).hexdigest()
    except OSError:
        pass
    return "unknown"


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "exact_scan.c")
    stamp_path = os.path.join(_HERE, "_build", "cpu.stamp")
    stamp = _cpu_stamp()
    try:
        same_cpu = open(stamp_path).read().strip() == stamp
    except OSError:
        same_cpu = False
    if (force or not same_cpu or not os.path.exists(_SO)
            or os.path.getmtime(_SO) < os.path.getmtime(src)):
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B"])
        with open(stamp_path, "w") as f:
            f.write(stamp)
    return _SO


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
        _lib.fro_exact_topk.restype = ctypes.c_int
        _lib.fro_exact_topk.argtypes = [
            ctypes.c_void_p, ctypes.c_int64, ctypes.c_int, ctypes.c_int, ctypes.c_void_p,
            ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p,
            ctypes.c_void_p, ctypes.c_int,
        ]
        _lib.fro_normalize_rows.restype = None
        _lib.fro_normalize_rows.argtypes = [ctypes.c_void_p, ctypes.c_int64, ctypes.c_int]
    return _lib


def exact_topk_prepared(corpus: np.ndarray, queries: np.ndarray, k: int, space: str = "cosine",
                        live: np.ndarray | None = None, nthreads: int = 0):
    """corpus: prepared [N,D] fp32, or uint16 raw-bf16; queries prepared [B,D] fp32."""
    assert corpus.flags.c_contiguous and queries.flags.c_contiguous
    n, dim = corpus.shape
    eb = 4 if corpus.dtype == np.float32 else 2
    assert corpus.dtype in (np.float32, np.uint16) and queries.dtype == np.float32
    b = queries.shape[0]
    out_d = np.empty((b, k), np.float32)
    out_r = np.empty((b, k), np.int64)
    lv = None
    if live is not None:
        lv = np.ascontiguousarray(live, dtype=np.uint8)
    t = lib().fro_exact_topk(
        corpus.ctypes.data, n, dim, eb, queries.ctypes.data, b, k, _SPACE[space],
        lv.ctypes.data if lv is not None else None, out_d.ctypes.data, out_r.ctypes.data, nthreads,
    )
    return out_d, out_r, t
