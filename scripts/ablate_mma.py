"""Ablation of the K2 large-batch scan: which of {corpus loads, accumulator reads} paces the tensor pipe?
(run on the GPU box; results of the debug modes are wrong by construction)"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import financial_rag_b200 as frb

n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
dev = torch.device("cuda", 0)
ix = frb.ShardIndex(dim=384, space="cosine", dtype="bf16", reserve_rows=n)
for c in range((n + 499_999) // 500_000):
    g = torch.Generator(device=dev).manual_seed(1234 + c)
    rows = min(500_000, n - c * 500_000)
    ix.append_device(torch.randn((rows, 384), generator=g, device=dev), None, first_key=c * 500_000)
torch.cuda.synchronize()
ix.set_path("mma")
k = 10
for b in (128, 256, 1024):
    q = torch.randn((b, 384), device=dev)
    od = torch.empty((b, k), dtype=torch.float32, device=dev)
    ok = torch.empty((b, k), dtype=torch.int64, device=dev)
    for co in (1, 2):
        if b <= 256 and co == 2:
            continue
        ix.set_option("mma_co_groups", co)
        for dbg in (0, 1, 2, 3):
            ix.set_option("mma_debug", dbg)
            for _ in range(3):
                ix.search_device(q, k, od, ok)
            torch.cuda.synchronize()
            ix.profile_read(); ix.set_profile(True)
            steps = 10
            for _ in range(steps):
                ix.search_device(q, k, od, ok)
            torch.cuda.synchronize()
            ix.set_profile(False)
            scan_ms, launches, _ = ix.profile_read()
            print(json.dumps({"batch": b, "co": co, "dbg": dbg, "scan_ms_per_step": round(scan_ms / steps, 4),
                              "tflops_scan": round(2.0 * n * 384 * b / (scan_ms / steps) / 1e9, 1),
                              "gbs_scan": round(n * 768 * launches / steps / (scan_ms / steps) / 1e6, 1)}), flush=True)
ix.set_option("mma_debug", 0)
