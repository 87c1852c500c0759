"""One search at batch 128 (K2, one CTA per SM) and 256 (K2, CTA pairs) over a 10M-row Gaussian corpus: launch set
for an ncu capture   ncu --set full --import-source on -k regex:scan_mma_kernel -o out python scripts/ncu_mid.py"""
import os, sys
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
for b in (128, 256):
    q = torch.randn((b, 384), generator=g, device=dev)
    od = torch.empty((b, 10), dtype=torch.float32, device=dev)
    ok = torch.empty((b, 10), dtype=torch.int64, device=dev)
    ix.search_device(q, 10, od, ok)
    torch.cuda.synchronize()
