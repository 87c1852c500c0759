"""What the reference's callers see: latency of store.search(vec, top_k) through get_child_vector_store (dict building,
payload lookup and all), one thread and several, on collections of the reference's size.
   python scripts/latency_store.py > gpurun_out/r02_latency_store_api.jsonl"""
import json, os, statistics, sys, tempfile, threading, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import financial_rag_b200 as frb


class Child:
    def __init__(self, cid, pid, content, emb):
        self.child_id, self.parent_id, self.content, self.embedding, self.context = cid, pid, content, emb, None


os.environ["B200_CHILD_AUTOPERSIST"] = "0"
devices = sys.argv[1] if len(sys.argv) > 1 else ""
if devices:
    os.environ["B200_CHILD_DEVICES"] = devices
for n in (1_000, 10_000, 100_000, 1_000_000):
    os.environ["CHROMA_CHILD_PERSIST_DIR"] = tempfile.mkdtemp()
    frb.reset_registry()
    rng = np.random.default_rng(n)
    store = frb.get_child_vector_store(collection="lat")
    for lo in range(0, n, 50_000):
        m = min(50_000, n - lo)
        vecs = rng.standard_normal((m, 384), dtype=np.float32)
        store.upsert_children([Child(10_000_000 + lo + i, (lo + i) // 4, f"snippet of child {lo + i} " * 8, vecs[i]) for i in range(m)])
    q = rng.standard_normal((64, 384), dtype=np.float32)
    for top_k in (6, 24):
        for _ in range(20):
            store.search(q[0].tolist(), top_k=top_k)
        ts = []
        for i in range(200):
            v = q[i % 64].tolist()
            t0 = time.perf_counter()
            hits = store.search(v, top_k=top_k)
            ts.append((time.perf_counter() - t0) * 1e6)
        assert len(hits) == top_k
        # four request threads, each with its own store object (rag_backend.py:611-643 constructs stores per request)
        def worker(out):
            st = frb.get_child_vector_store(collection="lat")
            t0 = time.perf_counter()
            for i in range(200):
                st.search(q[i % 64].tolist(), top_k=top_k)
            out.append(time.perf_counter() - t0)
        outs = []
        th = [threading.Thread(target=worker, args=(outs,)) for _ in range(4)]
        t0 = time.perf_counter()
        for t in th: t.start()
        for t in th: t.join()
        wall = time.perf_counter() - t0
        print(json.dumps({"children": n, "devices": devices or "0", "top_k": top_k, "search_us_median": round(statistics.median(ts), 1),
                          "search_us_p95": round(sorted(ts)[189], 1), "searches_per_s_1_thread": round(1e6 / statistics.median(ts)),
                          "searches_per_s_4_threads": round(800 / wall)}), flush=True)
frb.reset_registry()
