"""A/B of K2's shared score histogram (option mma_score_hist 1 / 0), scan-launch time per (rows, batch, k):
    python scripts/ab_k2_hist.py
FR_AB_MODE=slots: histogram on, threshold slots skipped once the histogram bounds the warp's queries (mma_debug 0) vs
always read (mma_debug 4096)."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import financial_rag_b200 as frb

dev = torch.device("cuda", 0)
cases = {390_625: [(4096, 10)], 10_000_000: [(1024, 10), (128, 10), (256, 10), (1024, 50), (1024, 100), (200, 50)],
         12_500_000: [(4096, 10)]}
sizes = sorted(cases)
ix = frb.ShardIndex(dim=384, dtype="bf16", reserve_rows=sizes[-1])
have = 0
for n in sizes:
    while have < n:
        rows = min(500_000, n - have)
        gg = torch.Generator(device=dev).manual_seed(1234 + have // 500_000)
        ix.append_device(torch.randn((rows, 384), generator=gg, device=dev), None, first_key=have)
        have += rows
    for batch, k in cases[n]:
        g = torch.Generator(device=dev).manual_seed(4321)
        q = torch.randn((batch, 384), generator=g, device=dev)
        ref = None
        for rep in range(2):
            for hist in (1, 0):
                if os.environ.get("FR_AB_MODE") == "slots":
                    ix.set_option("mma_debug", 0 if hist else 4096)
                elif os.environ.get("FR_AB_MODE") == "masks":  # pass masks built per group of 8 columns vs all 64
                    ix.set_option("mma_debug", 0 if hist else 8192)
                elif os.environ.get("FR_AB_MODE") == "refresh":  # geometric schedule up to every 64th tile vs every 8th
                    ix.set_option("mma_debug", 0 if hist else 16)
                else:
                    ix.set_option("mma_score_hist", hist)
                for _ in range(3):
                    d, kk = ix.search_device(q, k)
                torch.cuda.synchronize()
                if ref is None:
                    ref = kk.clone()
                same = bool((kk == ref).all())
                steps = max(4, min(40, int(4e9 / (n * max(batch, 256) / 2048 * 8))))
                ix.profile_read(); ix.set_profile(True)
                for _ in range(steps):
                    ix.search_device(q, k)
                torch.cuda.synchronize()
                ix.set_profile(False)
                ms, launches, searches = ix.profile_read()
                print(json.dumps({"rows": n, "batch": batch, "k": k, "hist": hist, "mode": os.environ.get("FR_AB_MODE", "hist"), "rep": rep, "ids_equal": same,
                                  "scan_ms_per_search": round(ms / searches, 4), "launches_per_search": launches / searches,
                                  "uncertified": ix.stat("mma_uncertified_queries")}), flush=True)
                time.sleep(0.2)
