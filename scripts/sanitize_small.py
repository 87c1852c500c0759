"""A small tour of the round-2 kernels for compute-sanitizer --tool memcheck (tiny sizes: the tool is 10-50x slower):
   compute-sanitizer --tool memcheck python scripts/sanitize_small.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import json
import numpy as np
import torch
import financial_rag_b200 as frb

rng = np.random.default_rng(0)
corpus = rng.standard_normal((3000, 384), dtype=np.float32)
keys = np.arange(3000, dtype=np.int64)
q = rng.standard_normal((70, 384), dtype=np.float32)
for space in ("cosine", "ip", "l2"):
    ix = frb.ShardIndex(dim=384, space=space, dtype="bf16")
    ix.upsert(corpus, keys)
    ix.set_path("mma")
    for b, k in ((1, 10), (20, 50), (64, 50), (30, 100), (70, 10)):
        d, kk = ix.search(q[:b], k)
        assert (kk[:, 0] >= 0).all()
    ix.set_path("stream")
    ix.search(q[:3], 10)
    ix.close()
for exchange in ("peer", "copy"):
    grp = frb.ShardGroup(dim=384, dtype="bf16", devices=[0, 0, 0], exchange=exchange)
    grp.upsert(corpus, keys)
    for b, k in ((1, 10), (33, 50), (70, 100)):
        grp.search(q[:b], k)
    grp.close()
dist = rng.random((3, 4, 20)).astype(np.float32)
kk = rng.integers(0, 30, size=(3, 4, 20)).astype(np.int64)
frb.score_fuse_host(dist, kk, 10)
frb.rrf_fuse_host(kk, 60, 10)
spec = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "encoder_models.json")))["thenlper/gte-small"]
from test_gpu_encoder import reference_model
cfg = dict(spec["config"]); cfg["num_hidden_layers"] = 2
enc = frb.B200QueryEncoder(cfg, reference_model(cfg, 1, layers=2).state_dict(), pooling="mean")
ids = rng.integers(0, 30522, size=(3, 9)).astype(np.int32)
for _ in range(3):
    enc.encode_ids(ids, np.array([9, 4, 1], np.int32))
enc.close()
print("sanitize tour ok")
