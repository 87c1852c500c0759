"""Host cost of enqueueing one sharded search from ONE process: the time fr_group_search_device takes to return (no
synchronise inside the timer), with the per-shard enqueueing threads on and off, and the host-buffer search latency.
   python scripts/enqueue_cost.py [shards] [rows_per_shard]        (shards on cuda:0..G-1 if there are G GPUs, else all on 0)"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import financial_rag_b200 as frb

shards = int(sys.argv[1]) if len(sys.argv) > 1 else 8
rows = int(sys.argv[2]) if len(sys.argv) > 2 else 200_000
ng = torch.cuda.device_count()
devices = [j % ng for j in range(shards)]
rng = np.random.default_rng(0)
grp = frb.ShardGroup(dim=384, dtype="bf16", devices=devices)
grp.upsert(rng.standard_normal((rows * shards, 384), dtype=np.float32), np.arange(rows * shards, dtype=np.int64))
out = {"shards": shards, "devices": devices, "rows_per_shard": rows, "exchange": grp.exchange}
for b in (1, 64):
    qh = rng.standard_normal((b, 384), dtype=np.float32)
    qd = [torch.from_numpy(qh).to(f"cuda:{d}") for d in devices]
    ref = None
    for threads in (1, 0, 1):
        grp.set_option("enqueue_threads", threads)
        for _ in range(20):
            grp.search_device(qd, 10, merge_on=[0])
        for d in set(devices):
            torch.cuda.synchronize(d)
        enq, tot = [], []
        for _ in range(300):
            t0 = time.perf_counter()
            d_, k_ = grp.search_device(qd, 10, merge_on=[0])
            t1 = time.perf_counter()
            for d in set(devices):
                torch.cuda.synchronize(d)
            t2 = time.perf_counter()
            enq.append(t1 - t0)
            tot.append(t2 - t0)
        keys = k_[0].cpu().numpy()
        ref = keys if ref is None else ref
        assert (keys == ref).all()
        host = []
        for _ in range(300):
            t0 = time.perf_counter()
            grp.search(qh, 10)
            host.append(time.perf_counter() - t0)
        out[f"b{b}_threads{threads}" + ("_again" if threads and f"b{b}_threads1" in out else "")] = {
            "enqueue_us": round(float(np.median(enq)) * 1e6, 1), "enqueue_plus_sync_us": round(float(np.median(tot)) * 1e6, 1),
            "host_search_us": round(float(np.median(host)) * 1e6, 1)}
print(json.dumps(out))
grp.close()
