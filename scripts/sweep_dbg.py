import json, os, sys
sys.path.insert(0, "/root/repo")
import torch
import financial_rag_b200 as frb
n = int(sys.argv[1]); cases=[int(x) for x in sys.argv[2].split(",")]; flags=[int(x) for x in sys.argv[3].split(",")]
dev = torch.device("cuda", 0)
ix = frb.ShardIndex(dim=384, space="cosine", dtype="bf16", reserve_rows=n)
for c in range((n + 499_999) // 500_000):
    g = torch.Generator(device=dev).manual_seed(1234 + c)
    rows = min(500_000, n - c * 500_000)
    ix.append_device(torch.randn((rows, 384), generator=g, device=dev), None, first_key=c * 500_000)
torch.cuda.synchronize()
ix.set_path("mma"); ix.set_option("mma_small_max", int(os.environ.get("SMALL_MAX", "0")))
k=10
for b in cases:
    q = torch.randn((b, 384), device=dev)
    od = torch.empty((b, k), dtype=torch.float32, device=dev); ok = torch.empty((b, k), dtype=torch.int64, device=dev)
    ref=None
    for rep in range(2):
        for f in flags:
            ix.set_option("mma_debug", f)
            steps = max(5, int(1e11 / (n * max(b, 128)) * 10))
            for _ in range(3): ix.search_device(q, k, od, ok)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(steps): ix.search_device(q, k, od, ok)
            e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / steps
            if ref is None: ref = ok.clone(); same=None
            else: same = bool((ref==ok).all())
            print(json.dumps({"rows": n, "batch": b, "dbg": f, "rep": rep, "ms": round(ms,4), "gbs": round(n*768/ms/1e6,1), "tflops": round(2.0*n*384*b/ms/1e9,1), "same": same}), flush=True)
