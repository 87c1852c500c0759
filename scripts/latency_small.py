"""Per-search latency on SMALL collections (the reference's real scale: 1e3 ... 1e6 child chunks), where the cost is
launch overhead, not bandwidth.  Host API (numpy in / numpy out, sync inside) and device API, per path.
   python scripts/latency_small.py"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import financial_rag_b200 as frb

dev = torch.device("cuda", 0)
k = 10
for n in (1_000, 10_000, 100_000, 1_000_000):
    ix = frb.ShardIndex(dim=384, space="cosine", dtype="bf16", reserve_rows=n)
    g = torch.Generator(device=dev).manual_seed(5)
    ix.append_device(torch.randn((n, 384), generator=g, device=dev), None, first_key=0)
    torch.cuda.synchronize()
    for b in (1, 8):
        qh = np.random.default_rng(b).standard_normal((b, 384)).astype(np.float32)
        qd = torch.from_numpy(qh).to(dev)
        od = torch.empty((b, k), dtype=torch.float32, device=dev)
        ok = torch.empty((b, k), dtype=torch.int64, device=dev)
        for path in ("stream", "mma", "auto"):
            ix.set_path(path)
            for _ in range(20):
                ix.search(qh, k)
            t0 = time.perf_counter()
            for _ in range(200):
                ix.search(qh, k)
            host_us = (time.perf_counter() - t0) / 200 * 1e6
            for _ in range(20):
                ix.search_device(qd, k, od, ok)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0 = time.perf_counter()
            e0.record()
            for _ in range(200):
                ix.search_device(qd, k, od, ok)
            e1.record(); torch.cuda.synchronize()
            wall_us = (time.perf_counter() - t0) / 200 * 1e6
            print(json.dumps({"rows": n, "batch": b, "path": path, "host_api_us": round(host_us, 1),
                              "device_api_gpu_us": round(e0.elapsed_time(e1) / 200 * 1e3, 1),
                              "device_api_wall_us": round(wall_us, 1)}), flush=True)
    ix.close()
