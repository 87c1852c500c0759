"""Stress of concurrent host searches on one shard (tests/test_gpu_parity.py::test_concurrent_host_searches_share_a_shard,
repeated on fresh shards so that slot growth, graph capture and replay keep meeting each other):
    python scripts/stress_concurrent.py [rounds]"""
import faulthandler, os, sys, threading
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import financial_rag_b200 as frb

if os.environ.get("FRB200_SEGV_TRACE") != "1":
    faulthandler.enable()
rounds = int(sys.argv[1]) if len(sys.argv) > 1 else 30
rng = np.random.default_rng(9100)
n = 60000
corpus = rng.standard_normal((n, 384), dtype=np.float32)
bad = 0
for r in range(rounds):
    ix = frb.ShardIndex(dim=384, space="cosine", dtype="bf16")
    ix.upsert(corpus, np.arange(n, dtype=np.int64) + 1000)
    if os.environ.get("FR_STRESS_NO_GRAPHS") == "1":
        ix.set_option("use_graphs", 0)
    ix.set_option("host_debug", int(os.environ.get("FR_STRESS_HOST_DEBUG", "0")))
    ixs = [ix]
    if os.environ.get("FR_STRESS_TWO") == "1":  # a second shard: graph work on one object beside waits on another
        jx = frb.ShardIndex(dim=384, space="cosine", dtype="bf16")
        jx.upsert(corpus[: n // 2], np.arange(n // 2, dtype=np.int64) + 1000)
        ixs.append(jx)
    jobs = []
    for t in range(7):
        b = (1, 3, 8, 40, 70, 200, 300)[t]
        q = rng.standard_normal((b, 384), dtype=np.float32)
        jobs.append((q, ixs[t % len(ixs)].search(q, 10)))
    errors = []

    def worker(t):
        q, (want_d, want_k) = jobs[t]
        try:
            for _ in range(25):
                d, kk = ixs[t % len(ixs)].search(q, 10)
                if not ((kk == want_k).all() and (d == want_d).all()):
                    errors.append(f"thread {t}: answer differs from the lone call")
                    return
        except Exception as e:  # noqa: BLE001
            errors.append(f"thread {t}: {e!r}")

    threads = [threading.Thread(target=worker, args=(t,)) for t in range(7)]
    for th in threads:
        th.start()
    for th in threads:
        th.join()
    if errors:
        bad += 1
        print(f"round {r}: {errors[:2]}", flush=True)
    for x in ixs:
        x.close()
print(f"{rounds} rounds, {bad} with errors", flush=True)
sys.exit(1 if bad else 0)
