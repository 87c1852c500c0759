"""ctypes binding of libfrb200.so -- the C ABI declared in include/fr_index.h.

This is the stub a maintainer of the reference would add next to
``parent_child/chroma_child_store.py``.  There is no CPU fallback: if the
shared library is missing or a call fails, an exception is raised.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, c_char_p, c_double, c_float, c_int, c_int64, c_uint64, c_void_p

_PKG = os.path.dirname(os.path.abspath(__file__))
# FRB200_LIB names another build of the same sources (scripts/build_variant.py: A/B runs of compile-time knobs);
# it is loaded and checked exactly like the default library
LIB_PATH = os.environ.get("FRB200_LIB") or os.path.join(_PKG, "libfrb200.so")

FR_COSINE, FR_L2, FR_IP = 0, 1, 2
FR_BF16, FR_F32 = 0, 1
FR_PATH_AUTO, FR_PATH_STREAM, FR_PATH_MMA = 0, 1, 2
FR_MAX_K = 128
FR_KEY_NONE = -1
FR_ABI_VERSION = 2
FR_XCHG_AUTO, FR_XCHG_NCCL, FR_XCHG_COPY, FR_XCHG_PEER = 0, 1, 2, 3

# every symbol include/fr_index.h declares: name -> (restype, argtypes)
SYMBOLS = {
    "fr_abi_version": (c_int, []),
    "fr_last_error": (c_char_p, []),
    "fr_launch_count": (c_int64, []),
    "fr_index_create": (c_int, [c_int, c_int, c_int, c_int, c_int64, POINTER(c_void_p)]),
    "fr_index_destroy": (c_int, [c_void_p]),
    "fr_index_reserve": (c_int, [c_void_p, c_int64]),
    "fr_index_set_option": (c_int, [c_void_p, c_char_p, c_int64]),
    "fr_index_count": (c_int, [c_void_p, POINTER(c_int64)]),
    "fr_index_rows": (c_int, [c_void_p, POINTER(c_int64)]),
    "fr_index_upsert": (c_int, [c_void_p, c_void_p, c_void_p, c_int64]),
    "fr_index_delete": (c_int, [c_void_p, c_void_p, c_int64, POINTER(c_int64)]),
    "fr_index_append_device": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_void_p]),
    "fr_index_get_rows": (c_int, [c_void_p, c_int64, c_int64, c_void_p, c_void_p]),
    "fr_index_export_raw": (c_int, [c_void_p, c_int64, c_int64, c_void_p, c_void_p]),
    "fr_index_import_raw": (c_int, [c_void_p, c_void_p, c_void_p, c_int64]),
    "fr_index_lookup_rows": (c_int, [c_void_p, c_void_p, c_int64, c_void_p]),
    "fr_index_get_stat": (c_int, [c_void_p, c_char_p, POINTER(c_int64)]),
    "fr_index_profile_read": (c_int, [c_void_p, POINTER(c_double), POINTER(c_int64), POINTER(c_int64)]),
    "fr_index_search": (c_int, [c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p]),
    "fr_index_search_device": (c_int, [c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "fr_index_search_partial_device": (c_int, [c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "fr_merge_shards_device": (c_int, [c_int, c_int, c_void_p, c_void_p, c_int64, c_int, c_int, c_int,
                                       c_void_p, c_void_p, c_void_p]),
    "fr_nccl_load": (c_int, [c_char_p]),
    "fr_nccl_version": (c_int, [POINTER(c_int)]),
    "fr_nccl_unique_id": (c_int, [c_void_p, c_int]),
    "fr_group_create": (c_int, [c_int, c_int, c_int, POINTER(c_int), c_int, c_int, c_int, c_void_p, c_int, c_int64,
                                POINTER(c_void_p)]),
    "fr_group_destroy": (c_int, [c_void_p]),
    "fr_group_info": (c_int, [c_void_p, c_char_p, POINTER(c_int64)]),
    "fr_group_set_option": (c_int, [c_void_p, c_char_p, c_int64]),
    "fr_group_reserve": (c_int, [c_void_p, c_int64]),
    "fr_group_shard": (c_int, [c_void_p, c_int, POINTER(c_void_p)]),
    "fr_group_adopt_rows": (c_int, [c_void_p, c_int64]),
    "fr_group_count": (c_int, [c_void_p, POINTER(c_int64)]),
    "fr_group_rows": (c_int, [c_void_p, POINTER(c_int64)]),
    "fr_group_upsert": (c_int, [c_void_p, c_void_p, c_void_p, c_int64]),
    "fr_group_delete": (c_int, [c_void_p, c_void_p, c_int64, POINTER(c_int64)]),
    "fr_group_get_rows": (c_int, [c_void_p, c_int64, c_int64, c_void_p, c_void_p]),
    "fr_group_export_raw": (c_int, [c_void_p, c_int64, c_int64, c_void_p, c_void_p]),
    "fr_group_import_raw": (c_int, [c_void_p, c_void_p, c_void_p, c_int64]),
    "fr_group_lookup_rows": (c_int, [c_void_p, c_void_p, c_int64, c_void_p]),
    "fr_group_search": (c_int, [c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p]),
    "fr_group_search_device": (c_int, [c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "fr_encoder_create": (c_int, [c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_float, POINTER(c_void_p)]),
    "fr_encoder_destroy": (c_int, [c_void_p]),
    "fr_encoder_set_tensor": (c_int, [c_void_p, c_char_p, c_void_p, c_int64, POINTER(c_int)]),
    "fr_encoder_finalize": (c_int, [c_void_p]),
    "fr_encoder_forward": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p]),
    "fr_encoder_forward_device": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p,
                                          c_void_p]),
    "fr_rrf_fuse": (c_int, [c_int, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p]),
    "fr_rrf_fuse_device": (c_int, [c_int, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p,
                                   c_void_p]),
    "fr_score_fuse": (c_int, [c_int, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p]),
    "fr_score_fuse_device": (c_int, [c_int, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "fr_maxsim_aggregate": (c_int, [c_int, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p]),
    "fr_maxsim_aggregate_device": (c_int, [c_int, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p,
                                           c_void_p, c_void_p]),
}

_lib = None


class FrError(RuntimeError):
    """A libfrb200 call returned a negative FR_E* code."""

    def __init__(self, code: int, message: str):
        super().__init__(f"libfrb200 error {code}: {message}")
        self.code = code


def load() -> ctypes.CDLL:
    """Load libfrb200.so (built in-tree by financial_rag_b200.build).  Fails loudly."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -m financial_rag_b200.build` "
            "(needs nvcc; the B200 backend has no CPU fallback)"
        )
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)  # AttributeError here = the header and the library disagree
        fn.restype = res
        fn.argtypes = args
    got = lib.fr_abi_version()
    if got != FR_ABI_VERSION:
        raise ImportError(f"libfrb200.so has ABI version {got}, binding expects {FR_ABI_VERSION}")
    _lib = lib
    return lib


_nccl_bound = False


def ensure_nccl() -> None:
    """Bind NCCL for the row-sharded groups (fr_group).  The library binds at run time (dlopen) so that the process
    holds ONE copy of NCCL: torch's bundled libnccl.so.2 when torch is (or will be) imported, else the system's."""
    global _nccl_bound
    if _nccl_bound:
        return
    import sys

    lib = load()
    path = None
    if "torch" not in sys.modules:  # with torch loaded, its copy is found by soname
        try:
            import importlib.util

            spec = importlib.util.find_spec("nvidia.nccl")
            for loc in (spec.submodule_search_locations if spec else []):
                cand = os.path.join(loc, "lib", "libnccl.so.2")
                if os.path.exists(cand):
                    path = cand
                    break
        except Exception:  # noqa: BLE001 - no bundled copy: the system's is next
            path = None
    check(lib.fr_nccl_load(path.encode() if path else None))
    _nccl_bound = True


def check(rc: int) -> None:
    if rc != 0:
        msg = load().fr_last_error()
        raise FrError(rc, msg.decode("utf-8", "replace") if msg else "unknown error")


def launch_count() -> int:
    return int(load().fr_launch_count())
