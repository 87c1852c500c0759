"""k = 100: 128 vs 256 candidates per query (option mma_wide_lists) on a Gaussian and a clustered corpus.
   python scripts/sweep_wide.py [rows] [batches] [k]"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import financial_rag_b200 as frb

n = int(sys.argv[1]) if len(sys.argv) > 1 else 40_000_000
batches = [int(x) for x in (sys.argv[2] if len(sys.argv) > 2 else "8,128,1024").split(",")]
k = int(sys.argv[3]) if len(sys.argv) > 3 else 100
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev).manual_seed(7)
centres = torch.nn.functional.normalize(torch.randn((2000, 384), generator=g, device=dev), dim=1)


def clustered(rows):
    which = torch.randint(0, 2000, (rows,), generator=g, device=dev)
    return centres[which] + 0.3 * torch.randn((rows, 384), generator=g, device=dev) / (384 ** 0.5)


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
        for wide in (0, 1):
            ix.set_option("mma_wide_lists", wide)
            for _ in range(2):
                ix.search_device(q, k, od, ok)
            torch.cuda.synchronize()
            u0, r0 = ix.stat("mma_uncertified_queries"), ix.stat("mma_rescanned_queries")
            steps = max(3, int(1e11 / (n * max(b, 128)) * 10))
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
                same = float((ref[0] - od).abs().max())
            print(json.dumps({"corpus": corpus, "rows": n, "batch": b, "k": k, "candidates": 256 if wide else 128,
                              "ms_per_search": round(ms, 3), "qps": round(b / ms * 1e3, 1),
                              "uncertified_per_search": round((ix.stat("mma_uncertified_queries") - u0) / steps, 1),
                              "rescanned_per_search": round((ix.stat("mma_rescanned_queries") - r0) / steps, 1),
                              "max_dist_diff_vs_128": same}), flush=True)
    ix.close()
    del ix
    torch.cuda.empty_cache()
