"""A batch-1024 search with k = 50 or 100 (k' = 128 / 256: K2 with its candidate lists in the partials slice) over 10M rows:
launch set for   ncu --set full --import-source on -k regex:scan_mma_kernel -s 2 -c 1 ...
   python scripts/ncu_wide_lists.py [k] [rows] [batch]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import financial_rag_b200 as frb

k = int(sys.argv[1]) if len(sys.argv) > 1 else 50
n = int(sys.argv[2]) if len(sys.argv) > 2 else 10_000_000
batch = int(sys.argv[3]) if len(sys.argv) > 3 else 1024
dev = torch.device("cuda", 0)
ix = frb.ShardIndex(dim=384, space="cosine", dtype="bf16", reserve_rows=n)
have = 0
while have < n:
    rows = min(500_000, n - have)
    gg = torch.Generator(device=dev).manual_seed(1234 + have // 500_000)
    ix.append_device(torch.randn((rows, 384), generator=gg, device=dev), None, first_key=have)
    have += rows
g = torch.Generator(device=dev).manual_seed(7)
q = torch.randn((batch, 384), generator=g, device=dev)
for _ in range(4):
    ix.search_device(q, k)
torch.cuda.synchronize()
ix.profile_read(); ix.set_profile(True)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    ix.search_device(q, k)
e1.record(); torch.cuda.synchronize()
ms, launches, searches = ix.profile_read()
print({"k": k, "rows": n, "batch": batch, "step_ms": e0.elapsed_time(e1) / 10, "scan_launch_ms": ms / max(launches, 1),
       "uncertified": ix.stat("mma_uncertified_queries")})
