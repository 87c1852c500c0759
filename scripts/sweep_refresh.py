"""K2: dense threshold refresh (every 8 tiles; diagnostics bit 16) vs the geometric schedule, same process, interleaved.
   python scripts/sweep_refresh.py [rows] [batch,k;batch,k;...]"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import financial_rag_b200 as frb

n = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000_000
cases = [tuple(int(v) for v in c.split(",")) for c in (sys.argv[2] if len(sys.argv) > 2 else "1024,10;1024,100;128,10;256,10").split(";")]
dev = torch.device("cuda", 0)
ix = frb.ShardIndex(dim=384, space="cosine", dtype="bf16", reserve_rows=n)
for c in range((n + 499_999) // 500_000):
    g = torch.Generator(device=dev).manual_seed(1234 + c)
    rows = min(500_000, n - c * 500_000)
    ix.append_device(torch.randn((rows, 384), generator=g, device=dev), None, first_key=c * 500_000)
torch.cuda.synchronize()
ix.set_path("mma")
for b, k in cases:
    q = torch.randn((b, 384), device=dev)
    od = torch.empty((b, k), dtype=torch.float32, device=dev)
    ok = torch.empty((b, k), dtype=torch.int64, device=dev)
    steps = max(3, int(2e11 / (n * max(b, 256)) * 20))
    ref = None
    for rep in range(2):
        for dense in (1, 0):
            ix.set_option("mma_debug", 16 if dense else 0)
            for _ in range(2):
                ix.search_device(q, k, od, ok)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(steps):
                ix.search_device(q, k, od, ok)
            e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / steps
            same = None
            if ref is None:
                ref = ok.clone()
            else:
                same = bool((ref == ok).all())
            print(json.dumps({"rows": n, "batch": b, "k": k, "refresh": "every 8 tiles" if dense else "geometric, cap 64",
                              "rep": rep, "ms_per_step": round(ms, 3), "qps": round(b / ms * 1e3, 1),
                              "tflops": round(2.0 * n * 384 * b / ms / 1e9, 1), "gbs": round(n * 768 / ms / 1e6, 1),
                              "ids_equal": same}), flush=True)
ix.set_option("mma_debug", 0)
