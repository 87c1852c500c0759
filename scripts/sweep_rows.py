"""Scan-launch time of the batched tensor-core scan as a function of the shard size (what does a launch cost beyond
its rows?).  Run on the GPU box:  python scripts/sweep_rows.py [batch] [co_groups,...] [lead,...]"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import financial_rag_b200 as frb

batch = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
cos = [int(x) for x in (sys.argv[2] if len(sys.argv) > 2 else "8").split(",")]
leads = [int(x) for x in (sys.argv[3] if len(sys.argv) > 3 else "6").split(",")]
dev = torch.device("cuda", 0)
sizes = [int(x) for x in os.environ.get("FR_SWEEP_SIZES", "1562500,3125000,6250000,12500000,25000000,50000000").split(",")]
dbgs = [int(x) for x in os.environ.get("FR_SWEEP_DBG", "0").split(",")]
ix = frb.ShardIndex(dim=384, dtype="bf16", reserve_rows=sizes[-1])
g = torch.Generator(device=dev).manual_seed(4321)
q = torch.randn((batch, 384), generator=g, device=dev)
have = 0
for n in sizes:
    while have < n:
        rows = min(500_000, n - have)
        gg = torch.Generator(device=dev).manual_seed(1234 + have // 500_000)
        ix.append_device(torch.randn((rows, 384), generator=gg, device=dev), None, first_key=have)
        have += rows
    torch.cuda.synchronize()
    for co in cos:
      for dbg in dbgs:
        for lead in leads:
            ix.set_option("mma_debug", dbg)
            ix.set_option("mma_co_groups", co)
            ix.set_option("mma_max_lead", lead)
            for _ in range(3):
                ix.search_device(q, 10)
            torch.cuda.synchronize()
            steps = max(4, min(40, int(2e9 / (n * batch / 2048 * 8))))
            ix.profile_read(); ix.set_profile(True)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(steps):
                ix.search_device(q, 10)
            e1.record(); torch.cuda.synchronize()
            ix.set_profile(False)
            ms, launches, searches = ix.profile_read()
            print(json.dumps({"rows": n, "batch": batch, "co": co, "lead": lead, "dbg": dbg, "steps": steps,
                              "step_ms": round(e0.elapsed_time(e1) / steps, 4), "scan_launch_ms": round(ms / launches, 4),
                              "launches_per_step": launches / steps,
                              "tflops": round(2.0 * n * 384 * batch / (e0.elapsed_time(e1) / steps) / 1e9, 1),
                              "ns_per_row_launch": round(ms / launches * 1e6 / n, 4)}), flush=True)
            time.sleep(0.3)
