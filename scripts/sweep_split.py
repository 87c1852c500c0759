"""Cost of two-term (hi + lo) bf16 queries in K2s, and what they buy on a clustered corpus (run on the GPU box).
For each corpus (isotropic Gaussian, 2000 clusters) and batch: mma_split 0 / 1 -> ms per search, scan GB/s,
first-pass certification failures per search.
   python scripts/sweep_split.py [rows] [batches]"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import financial_rag_b200 as frb

n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
batches = [int(x) for x in (sys.argv[2] if len(sys.argv) > 2 else "1,4,8,16,32,64").split(",")]
k = 10
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev).manual_seed(7)
centres = torch.nn.functional.normalize(torch.randn((2000, 384), generator=g, device=dev), dim=1)
spread = 0.3


def clustered(rows):
    which = torch.randint(0, 2000, (rows,), generator=g, device=dev)
    return centres[which] + spread * torch.randn((rows, 384), generator=g, device=dev) / (384 ** 0.5)


for corpus in ("gaussian", "clustered"):
    ix = frb.ShardIndex(dim=384, space="cosine", dtype="bf16", reserve_rows=n)
    for c in range((n + 499_999) // 500_000):
        rows = min(500_000, n - c * 500_000)
        x = torch.randn((rows, 384), generator=g, device=dev) if corpus == "gaussian" else clustered(rows)
        ix.append_device(x, None, first_key=c * 500_000)
    torch.cuda.synchronize()
    ix.set_path("mma")
    for b in batches:
        q = torch.randn((b, 384), generator=g, device=dev) if corpus == "gaussian" else clustered(b)
        od = torch.empty((b, k), dtype=torch.float32, device=dev)
        ok = torch.empty((b, k), dtype=torch.int64, device=dev)
        ref = None
        for split in (0, 1):
            ix.set_option("mma_split", split)
            for _ in range(3):
                ix.search_device(q, k, od, ok)
            torch.cuda.synchronize()
            u0, r0 = ix.stat("mma_uncertified_queries"), ix.stat("mma_rescanned_queries")
            steps = 20
            ix.profile_read(); ix.set_profile(True)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(steps):
                ix.search_device(q, k, od, ok)
            e1.record(); torch.cuda.synchronize()
            ix.set_profile(False)
            scan_ms, launches, _ = ix.profile_read()
            ms = e0.elapsed_time(e1) / steps
            same = None
            if ref is None:
                ref = ok.clone()
            else:
                same = bool((ref == ok).all())
            print(json.dumps({"corpus": corpus, "rows": n, "batch": b, "split": split, "ms_per_search": round(ms, 4),
                              "qps": round(b / ms * 1e3, 1), "first_pass_gbs": round(n * 768 / (scan_ms / launches) / 1e6, 1),
                              "uncertified_per_search": (ix.stat("mma_uncertified_queries") - u0) / steps,
                              "rescanned_per_search": (ix.stat("mma_rescanned_queries") - r0) / steps,
                              "ids_equal_split0": same}), flush=True)
    ix.close()
    del ix
    torch.cuda.empty_cache()
