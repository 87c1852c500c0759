"""cfg3 of BASELINE.json: dual-encoder ensemble (2 collections x N rows x 384 bf16), per-collection top-50,
RRF (k = 60) to top-10, everything on the device (run on the GPU box).
   python scripts/bench_cfg3.py [rows_per_collection] [batches]"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import financial_rag_b200 as frb

n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
batches = [int(x) for x in (sys.argv[2] if len(sys.argv) > 2 else "1,64,1024").split(",")]
kp, k_out, k_rrf = 50, 10, 60
dev = torch.device("cuda", 0)
cols = []
for seed in (1234, 2234):  # the two collections use seeds 1234 / 2234
    ix = frb.ShardIndex(dim=384, space="cosine", dtype="bf16", reserve_rows=n)
    for c in range((n + 499_999) // 500_000):
        g = torch.Generator(device=dev).manual_seed(seed + c)
        rows = min(500_000, n - c * 500_000)
        ix.append_device(torch.randn((rows, 384), generator=g, device=dev), None, first_key=c * 500_000)
    cols.append(ix)
torch.cuda.synchronize()
for b in batches:
    g = torch.Generator(device=dev).manual_seed(4321)
    qs = [torch.randn((b, 384), generator=g, device=dev) for _ in cols]  # one embedding per encoder
    keys = torch.empty((2, b, kp), dtype=torch.int64, device=dev)
    dist = torch.empty((2, b, kp), dtype=torch.float32, device=dev)

    def step():
        for i, ix in enumerate(cols):
            ix.search_device(qs[i], kp, dist[i], keys[i])
        return frb.rrf_fuse_device(keys, k_rrf, k_out)

    for _ in range(3):
        step()
    torch.cuda.synchronize()
    steps = 10 if n * b < 4e10 else 4
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        sc, fused = step()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    print(json.dumps({"config": "cfg3 dual-encoder ensemble", "rows_per_collection": n, "batch": b, "k_per_collection": kp,
                      "k_out": k_out, "ms_per_step": round(ms, 4), "qps": round(b / ms * 1e3, 1),
                      "scan_gbs_equiv": round(2 * n * 768 / ms / 1e6, 1),
                      "uncertified": [ix.stat("mma_uncertified_queries") for ix in cols],
                      "rescanned": [ix.stat("mma_rescanned_queries") for ix in cols]}), flush=True)
