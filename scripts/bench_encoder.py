"""Latency of the query-side encoder (row f-4) on the GPU box: token ids -> normalised query block, per (batch, tokens),
next to transformers.BertModel on the host cores (what the reference's embedder does per query).
   python scripts/bench_encoder.py > gpurun_out/r02_encoder_latency.jsonl"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
import financial_rag_b200 as frb
from financial_rag_b200 import _lib
from test_gpu_encoder import reference_model

spec = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "encoder_models.json")))["BAAI/bge-small-en-v1.5"]
cfg = spec["config"]
model = reference_model(cfg, seed=1)
enc = frb.B200QueryEncoder(cfg, model.state_dict(), pooling=spec["pooling"])
dev = torch.device("cuda", 0)
torch.set_num_threads(os.cpu_count() or 1)
for B, T in ((1, 16), (1, 32), (8, 32), (64, 32), (64, 128), (16, 512)):
    ids = torch.randint(0, cfg["vocab_size"], (B, T), dtype=torch.int32, device=dev)
    lens = torch.full((B,), T, dtype=torch.int32, device=dev)
    out = torch.empty((B, 384), dtype=torch.float32, device=dev)
    for _ in range(5):
        enc.encode_ids_device(ids, lens, out)
    torch.cuda.synchronize()
    n = 50
    l0 = _lib.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for _ in range(n):
        enc.encode_ids_device(ids, lens, out)
    e1.record(); torch.cuda.synchronize()
    wall = (time.perf_counter() - t0) / n * 1e3
    gpu_ms = e0.elapsed_time(e1) / n
    launches = (_lib.launch_count() - l0) // n
    # host form (ids on the host, embedding back on the host)
    ids_h, lens_h = ids.cpu().numpy(), lens.cpu().numpy()
    t0 = time.perf_counter()
    for _ in range(20):
        enc.encode_ids(ids_h, lens_h)
    host_ms = (time.perf_counter() - t0) / 20 * 1e3
    cpu_ms = None
    if B * T <= 2048:
        ids64 = ids.cpu().to(torch.int64)
        with torch.no_grad():
            model(input_ids=ids64)
            t0 = time.perf_counter()
            for _ in range(5):
                model(input_ids=ids64)
            cpu_ms = (time.perf_counter() - t0) / 5 * 1e3
    flops = 12 * (2 * B * T * (4 * 384 * 384 + 2 * 384 * 1536) + 4 * B * 12 * T * T * 32)
    print(json.dumps({"batch": B, "tokens": T, "gpu_ms_device_api": round(gpu_ms, 4), "wall_ms_device_api": round(wall, 4),
                      "ms_host_api": round(host_ms, 4), "launches": int(launches), "model_gflop": round(flops / 1e9, 3),
                      "cpu_transformers_ms": None if cpu_ms is None else round(cpu_ms, 2), "cpu_threads": torch.get_num_threads()}), flush=True)
