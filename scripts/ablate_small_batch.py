"""Which part of K2 costs HBM bandwidth at small batches under the power cap?  Long runs (>= 2 s each) of
batch 8 / 64 over a 40M-row shard with: everything, no accumulator reads (2), no MMAs (4), neither (6).
(run on the GPU box; results of the debug modes are wrong by construction)"""
import json, os, sys, subprocess, threading, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import financial_rag_b200 as frb

n = int(sys.argv[1]) if len(sys.argv) > 1 else 40_000_000
dev = torch.device("cuda", 0)
ix = frb.ShardIndex(dim=384, space="cosine", dtype="bf16", reserve_rows=n)
for c in range((n + 499_999) // 500_000):
    g = torch.Generator(device=dev).manual_seed(1234 + c)
    rows = min(500_000, n - c * 500_000)
    ix.append_device(torch.randn((rows, 384), generator=g, device=dev), None, first_key=c * 500_000)
torch.cuda.synchronize()
k = 10


def clocks():
    out = subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,power.draw", "--format=csv,noheader,nounits", "-i", "0"],
                         capture_output=True, text=True).stdout.strip().split(",")
    return float(out[0]), float(out[1])


for path, b in (("stream", 1), ("mma", 8), ("mma", 64), ("mma", 128)):
    ix.set_path(path)
    q = torch.randn((b, 384), device=dev)
    od = torch.empty((b, k), dtype=torch.float32, device=dev)
    ok = torch.empty((b, k), dtype=torch.int64, device=dev)
    for dbg in ((0,) if path == "stream" else (0, 2, 4, 6)):
        ix.set_option("mma_debug", dbg)
        for _ in range(3):
            ix.search_device(q, k, od, ok)
        torch.cuda.synchronize()
        ix.profile_read(); ix.set_profile(True)
        steps = 400
        samples = []
        for i in range(steps):
            ix.search_device(q, k, od, ok)
            if i % 100 == 99:
                samples.append(clocks())
        torch.cuda.synchronize()
        ix.set_profile(False)
        scan_ms, launches, _ = ix.profile_read()
        print(json.dumps({"path": path, "batch": b, "dbg": dbg, "scan_ms": round(scan_ms / launches, 4),
                          "gbs": round(n * 768 / (scan_ms / launches) / 1e6, 1),
                          "sm_mhz_power_w": samples}), flush=True)
ix.set_option("mma_debug", 0)
