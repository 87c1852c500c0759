"""Sweep query batch sizes on one GPU: QPS, ms/step and scan-kernel GB/s per path (run on the GPU box)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import financial_rag_b200 as frb

n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
batches = [int(x) for x in (sys.argv[2] if len(sys.argv) > 2 else "1,2,4,8,16,32,64,128,256,1024").split(",")]
k = int(sys.argv[3]) if len(sys.argv) > 3 else 10
dev = torch.device("cuda", 0)
ix = frb.ShardIndex(dim=384, space="cosine", dtype="bf16", reserve_rows=n)
for c in range((n + 499_999) // 500_000):
    g = torch.Generator(device=dev).manual_seed(1234 + c)
    rows = min(500_000, n - c * 500_000)
    ix.append_device(torch.randn((rows, 384), generator=g, device=dev), None, first_key=c * 500_000)
torch.cuda.synchronize()
out = []
for path in (os.environ.get("SWEEP_PATHS", "stream,mma").split(",")):
    ix.set_path(path)
    for b in batches:
        if path == "stream" and b > 16:
            continue
        q = torch.randn((b, 384), device=dev)
        od = torch.empty((b, k), dtype=torch.float32, device=dev)
        ok = torch.empty((b, k), dtype=torch.int64, device=dev)
        steps = 10 if n * b < 4e10 else 3
        for _ in range(3):
            ix.search_device(q, k, od, ok)
        torch.cuda.synchronize()
        ix.profile_read(); ix.set_profile(True)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            ix.search_device(q, k, od, ok)
        e1.record(); torch.cuda.synchronize()
        ix.set_profile(False)
        scan_ms, launches, _ = ix.profile_read()
        ms = e0.elapsed_time(e1) / steps
        r = {"path": path, "rows": n, "batch": b, "k": k, "ms_per_step": round(ms, 4), "qps": round(b / ms * 1e3, 1),
             "scan_ms_per_launch": round(scan_ms / launches, 4), "launches_per_step": launches // steps,
             "scan_gbs": round(n * 768 / (scan_ms / launches) / 1e6, 1),
             "tflops": round(2.0 * n * 384 * b / ms / 1e9, 1), "uncertified_total": ix.stat("mma_uncertified_queries")}
        out.append(r)
        print(json.dumps(r), flush=True)
