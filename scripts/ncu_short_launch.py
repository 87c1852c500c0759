"""A batch-2048 search (K2, CTA pairs, 8 co-resident groups) over a SHORT shard, where the per-launch fixed cost of the
epilogue is most of the launch: launch set for   ncu --set full --import-source on -k regex:scan_mma_kernel -s 2 -c 1 ...
   python scripts/ncu_short_launch.py [rows]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import financial_rag_b200 as frb

n = int(sys.argv[1]) if len(sys.argv) > 1 else 390_625
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev).manual_seed(7)
ix = frb.ShardIndex(dim=384, space="cosine", dtype="bf16", reserve_rows=n)
ix.append_device(torch.randn((n, 384), generator=g, device=dev), None, first_key=0)
ix.set_path("mma")
q = torch.randn((2048, 384), generator=g, device=dev)
for _ in range(4):
    ix.search_device(q, 10)
torch.cuda.synchronize()
print("ok")
