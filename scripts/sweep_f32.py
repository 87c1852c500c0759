"""fp32 collection: batched search on the tensor cores (bf16 selection copy + exact fp32 rescoring) vs the CUDA-core
stream kernel (ceil(B/4) corpus passes).   python scripts/sweep_f32.py [rows] [batches]"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import financial_rag_b200 as frb

n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
batches = [int(x) for x in (sys.argv[2] if len(sys.argv) > 2 else "1,8,64,1024").split(",")]
k = 10
dev = torch.device("cuda", 0)
ix = frb.ShardIndex(dim=384, space="cosine", dtype="f32", reserve_rows=n)
for c in range((n + 499_999) // 500_000):
    g = torch.Generator(device=dev).manual_seed(1234 + c)
    rows = min(500_000, n - c * 500_000)
    ix.append_device(torch.randn((rows, 384), generator=g, device=dev), None, first_key=c * 500_000)
torch.cuda.synchronize()
for b in batches:
    q = torch.randn((b, 384), device=dev)
    od = torch.empty((b, k), dtype=torch.float32, device=dev)
    ok = torch.empty((b, k), dtype=torch.int64, device=dev)
    ref = None
    for path in ("stream", "mma"):
        if path == "stream" and b > 64:
            continue
        ix.set_path(path)
        steps = 10 if path == "mma" else 3
        for _ in range(2):
            ix.search_device(q, k, od, ok)
        torch.cuda.synchronize()
        u0 = ix.stat("mma_uncertified_queries")
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            ix.search_device(q, k, od, ok)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        same = None
        if ref is None:
            ref = (od.clone(), ok.clone())
        else:
            same = {"ids_equal": bool((ref[1] == ok).all()), "max_dist_diff": float((ref[0] - od).abs().max())}
        print(json.dumps({"rows": n, "storage": "f32", "batch": b, "path": path, "ms_per_search": round(ms, 3),
                          "qps": round(b / ms * 1e3, 1), "uncertified_per_search": (ix.stat("mma_uncertified_queries") - u0) / steps,
                          "vs_stream": same}), flush=True)
