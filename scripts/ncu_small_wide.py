"""K2s with wide candidate lists (k' = 128 / 256, lists shared by the four lane-quarter warps in shared memory) on a 10M-row
Gaussian corpus: search time per (batch, k) with CUDA events, or -- with FR_ONE=batch,k -- just three searches of one
case, the launch set for
   ncu --set full --clock-control none --import-source on -k regex:scan_mma_small -s 2 -c 1 -o out python scripts/ncu_small_wide.py"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import financial_rag_b200 as frb

n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev).manual_seed(7)
ix = frb.ShardIndex(dim=384, space="cosine", dtype="bf16", reserve_rows=n)
for c in range((n + 499_999) // 500_000):
    rows = min(500_000, n - c * 500_000)
    ix.append_device(torch.randn((rows, 384), generator=g, device=dev), None, first_key=c * 500_000)
ix.set_path("mma")
one = os.environ.get("FR_ONE")
cases = [tuple(int(x) for x in one.split(","))] if one else [(64, 10, 0), (64, 32, 0), (64, 50, 0), (64, 50, 2), (32, 50, 0), (16, 50, 0), (16, 10, 0),
                                                              (1, 10, 0), (1, 50, 0), (32, 100, 0), (16, 100, 0)]
hists = [int(x) for x in os.environ.get("FR_HIST", "1,0").split(",")]   # 0: threshold slots only, first tile inserted at once
for b, k, *rest in [c + (h,) if len(c) == 3 else c + (0, h) for c in cases for h in hists]:
    dbg, hist = rest[0], rest[1]
    ix.set_option("mma_score_hist", hist)
    ix.set_option("mma_debug", dbg)   # 2: gate only, nothing enters a list; 64: no threshold refresh (results wrong either way); 128: threshold slots start
    # from the previous (identical) search's final values -- what near-perfect sharing from the first tile would give
    q = torch.randn((b, 384), generator=g, device=dev)
    od = torch.empty((b, k), dtype=torch.float32, device=dev)
    ok = torch.empty((b, k), dtype=torch.int64, device=dev)
    for _ in range(3):
        ix.search_device(q, k, od, ok)
    torch.cuda.synchronize()
    if one:
        break
    ix.profile_read(); ix.set_profile(True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(30):
        ix.search_device(q, k, od, ok)
    e1.record(); torch.cuda.synchronize()
    ix.set_profile(False)
    ms, launches, searches = ix.profile_read()
    print(json.dumps({"hist": hist, "rows": n, "batch": b, "k": k, "dbg": dbg, "search_ms": round(e0.elapsed_time(e1) / 30, 4),
                      "scan_launch_ms": round(ms / max(launches, 1), 4), "hbm_floor_ms": round(n * 768 / 6547.2e6, 4)}), flush=True)
