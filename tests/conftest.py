import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden():
    """The 18 fp32 vectors of the reference's Chroma WAL (tests/golden/make_golden.py)."""
    with open(os.path.join(ROOT, "tests", "golden", "chroma_fixture.json")) as f:
        raw = json.load(f)
    cols = {}
    for r in raw["rows"]:
        c = cols.setdefault(r["collection"], {"ids": [], "vectors": [], "metadatas": [], "blobs": []})
        blob = bytes.fromhex(r["vector_f32le_hex"])
        c["ids"].append(r["id"])
        c["vectors"].append(np.frombuffer(blob, dtype="<f4").copy())
        c["metadatas"].append(r["metadata"])
        c["blobs"].append(blob)
    for c in cols.values():
        c["vectors"] = np.stack(c["vectors"]).astype(np.float32)
    return {"raw": raw, "collections": cols}


@pytest.fixture(scope="session")
def rrf_traces():
    with open(os.path.join(ROOT, "tests", "golden", "rrf_traces.json")) as f:
        return json.load(f)["traces"]
