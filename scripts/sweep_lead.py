"""K2 with co-resident query groups: how far a group may run ahead of the slowest group of its corpus stream
(option mma_max_lead; 0 = unthrottled) vs QPS, same process, interleaved repetitions.
   python scripts/sweep_lead.py [rows] [batch] [leads] [co_groups]"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import financial_rag_b200 as frb

n = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000_000
b = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
leads = [int(x) for x in (sys.argv[3] if len(sys.argv) > 3 else "0,1,3,8").split(",")]
cos = [int(x) for x in (sys.argv[4] if len(sys.argv) > 4 else "4").split(",")]
k = 10
dev = torch.device("cuda", 0)
ix = frb.ShardIndex(dim=384, space="cosine", dtype="bf16", reserve_rows=n)
for c in range((n + 499_999) // 500_000):
    g = torch.Generator(device=dev).manual_seed(1234 + c)
    rows = min(500_000, n - c * 500_000)
    ix.append_device(torch.randn((rows, 384), generator=g, device=dev), None, first_key=c * 500_000)
torch.cuda.synchronize()
ix.set_path("mma")
q = torch.randn((b, 384), device=dev)
od = torch.empty((b, k), dtype=torch.float32, device=dev)
ok = torch.empty((b, k), dtype=torch.int64, device=dev)
ref = None
steps = max(3, int(2e11 / (n * b) * 20))
for rep in range(2):
  for co in cos:
    ix.set_option("mma_co_groups", co)
    for lead in leads:
        ix.set_option("mma_max_lead", lead)
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
        print(json.dumps({"rows": n, "batch": b, "co_groups": co, "max_lead": lead, "rep": rep, "steps": steps, "ms_per_step": round(ms, 3),
                          "qps": round(b / ms * 1e3, 1), "tflops": round(2.0 * n * 384 * b / ms / 1e9, 1),
                          "ids_equal": same}), flush=True)
