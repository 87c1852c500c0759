"""Diagnostic: where does a device-path search step spend its time? (run on the GPU box)"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import financial_rag_b200 as frb
from financial_rag_b200.sharded import ShardedSearcher

n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
dev = torch.device("cuda", 0)
ix = frb.ShardIndex(dim=384, space="cosine", dtype="bf16", reserve_rows=n)
for c in range(n // 500_000):
    g = torch.Generator(device=dev).manual_seed(1234 + c)
    ix.append_device(torch.randn((500_000, 384), generator=g, device=dev), None, first_key=c * 500_000)
torch.cuda.synchronize()


def timeit(fn, steps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    t_host = (time.perf_counter() - t0) / steps * 1e3
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps, t_host


for b in (1, 4):
    q = torch.randn((b, 384), device=dev)
    k = 10
    s = ShardedSearcher(ix, k, b, device=dev)
    od = torch.empty((b, k), dtype=torch.float32, device=dev)
    ok = torch.empty((b, k), dtype=torch.int64, device=dev)
    print(f"B={b} n={n}")
    print("  search_device (final)      gpu/host ms:", timeit(lambda: ix.search_device(q, k, od, ok)))
    print("  search_partial_device      gpu/host ms:", timeit(lambda: ix.search_partial_device(q, k, s.local[0], s.local[1])))
    print("  sharded.search_device      gpu/host ms:", timeit(lambda: s.search_device(q)))
    ix.set_profile(True)
    r = timeit(lambda: s.search_device(q))
    ix.set_profile(False)
    print("  sharded + profile          gpu/host ms:", r, "scan:", ix.profile_read())
    qp = torch.empty((b, 384)).pin_memory(); qp.copy_(q.cpu())
    odp = torch.empty((b, k)).pin_memory(); okp = torch.empty((b, k), dtype=torch.int64).pin_memory()
    print("  host fr_index_search       gpu/host ms:", timeit(lambda: ix.search_raw(qp.data_ptr(), b, k, odp.data_ptr(), okp.data_ptr())))
