"""CPU-side checks of the boundary: the C ABI library loads and exports every symbol that
include/fr_index.h declares, the host mirror behaves like the reference's store up to the point
where a GPU is needed, and that point fails loudly (no CPU fallback)."""
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from financial_rag_b200 import build as frbuild

    frbuild.build()
    from financial_rag_b200 import _lib

    return _lib


def test_library_exports_every_declared_symbol(lib):
    header = open(os.path.join(ROOT, "include", "fr_index.h")).read()
    declared = set(re.findall(r"\b(fr_[a-z_0-9]+)\s*\(", header))
    assert declared, "no declarations parsed"
    assert declared == set(lib.SYMBOLS), declared ^ set(lib.SYMBOLS)
    cdll = lib.load()
    for name in declared:
        assert getattr(cdll, name) is not None
    assert cdll.fr_abi_version() == lib.FR_ABI_VERSION
    m = re.search(r"#define FR_MAX_K (\d+)", header)
    assert int(m.group(1)) == lib.FR_MAX_K


def test_built_for_sm_100a_only(lib):
    import subprocess

    out = subprocess.run(["cuobjdump", "-lelf", lib.LIB_PATH], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs


def test_no_gpu_means_loud_failure_not_fallback(lib):
    import torch

    if torch.cuda.is_available():
        pytest.skip("box has a GPU")
    from financial_rag_b200 import ShardIndex

    with pytest.raises(lib.FrError) as ei:
        ShardIndex(dim=384)
    assert ei.value.code == -4 and "no CPU fallback" in str(ei.value)
    with pytest.raises(lib.FrError):
        from financial_rag_b200 import rrf_fuse_host

        rrf_fuse_host(np.zeros((2, 1, 3), np.int64))


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "financial_rag_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, re.M), f
                assert "oracle/" not in src or f.endswith(".md"), f


def test_store_mirror_host_logic(lib, monkeypatch, tmp_path):
    """Constructor / env resolution / empty-collection behaviour of the ChromaChildStore mirror
    (parent_child/chroma_child_store.py:18-34, 76-80) -- no device needed until data arrives."""
    import financial_rag_b200 as frb

    frb.reset_registry()
    monkeypatch.delenv("CHROMA_CHILD_PERSIST_DIR", raising=False)
    monkeypatch.delenv("CHILD_VECTOR_COLLECTION", raising=False)
    s = frb.B200ChildStore()
    assert s.collection_name == "parent_child_children"
    assert s.persist_dir == os.path.join(ROOT, ".chroma_children")
    monkeypatch.setenv("CHROMA_CHILD_PERSIST_DIR", str(tmp_path))
    monkeypatch.setenv("CHILD_VECTOR_COLLECTION", "from_env")
    s = frb.B200ChildStore()
    assert (s.persist_dir, s.collection_name) == (str(tmp_path), "from_env")
    s = frb.get_child_vector_store(collection="children_x", table="ignored")
    assert s.collection_name == "children_x" and s.count() == 0
    assert s.col is frb.get_child_vector_store(collection="children_x").col  # process-global registry
    assert s.col.metadata == {"hnsw:space": "cosine"}
    assert s.search([0.0] * 384, top_k=6) == []
    assert s.upsert_children([]) is True

    class C:  # child without an embedding is skipped (chroma_child_store.py:41)
        child_id, parent_id, content, embedding, context = 1, 2, "x", None, None

    assert s.upsert_children([C()]) is True and s.count() == 0
    monkeypatch.setenv("CHILD_VECTOR_BACKEND", "nonsense")
    with pytest.raises(ValueError):
        frb.get_child_vector_store()
    frb.reset_registry()


def test_id_to_key_mapping():
    from financial_rag_b200.collection import B200Collection

    c = B200Collection("t", {"hnsw:space": "l2"})
    assert c.space == "l2"
    assert c._key_for("217959081514635264") == 217959081514635264
    a, b = c._key_for("217959081514635264:0"), c._key_for("abc")
    assert a == -2 and b == -3 and c._key_for("abc") == -3
    assert c._key_for("007") < 0  # not canonical decimal -> synthetic key
    assert c._key_for(str(1 << 63)) < 0  # does not fit int64
    from financial_rag_b200 import canonical_space

    assert [canonical_space(x) for x in ("cos", "Cosine", "euclidean", "l2", "ip", "inner_product", None, "zzz")] == \
        ["cosine", "cosine", "l2", "l2", "ip", "ip", "cosine", "cosine"]


def test_bench_reference_arm_prints_one_json_line():
    """bench.py --impl reference (the CPU arm the driver runs beside ours) needs no GPU and must print exactly
    one JSON line on stdout with the contract's keys."""
    import json
    import subprocess
    import sys

    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--rows", "300000",
                        "--batch", "4", "--steps", "1", "--warmup", "1"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, r.stdout
    d = json.loads(lines[0])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better",
                "scaling", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in d, key
    assert d["impl"] == "reference" and d["unit"] == "queries/s" and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
