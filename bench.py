#!/usr/bin/env python3
"""bench.py -- QPS of exact top-k over a synthetic unit-norm 384-d corpus (BASELINE.json; default cfg4: 100M x 384 bf16).

    python bench.py --gpus 1 --steps 50 --warmup 5                       # this repo's CUDA path, one GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...   # one process per GPU
    python bench.py --gpus N --single-process                            # one process owning N GPUs (the store's mode)
    python bench.py --config cfg2|cfg3|cfg5 ...                          # the other BASELINE.json configs
    python bench.py --impl reference ...                                 # CPU arm (see below)

A "step" is one search of one query batch over the whole corpus.  The corpus (fixed total size, row-sharded over the N
GPUs => "strong" scaling) is generated on the devices, chunk by chunk, from fixed seeds, so every N sees the same rows; it is
far larger than L2 (>= 9.6 GB per GPU), so no flush is needed between timed iterations.

What is measured is the product path: the row-sharded collection of the C ABI (fr_group: cyclic row placement, scan +
fused top-k per GPU, ONE NCCL all-gather of the local lists issued by the library itself, merge kernel) -- the same object
``get_child_vector_store`` hands out under ``B200_CHILD_DEVICES``; with one GPU it is the plain fr_index.

  value     whole-job QPS with the query block already resident in HBM (CUDA events on the launching streams, max over ranks)
  e2e       the same through the host-buffer C-ABI call (fr_index_search / fr_group_search: pinned query block ->
            H2D [-> ncclBroadcast] -> scan -> all-gather -> merge -> D2H), every sweep entry carries its own
  roofline  the scan kernel alone: algorithmic bytes (rows_per_gpu * 768) / its mean launch time, bracketed by CUDA events
            inside the library on the launching stream
  parity    the literal north-star gate on a 64-query block of the timed batch: ids / scores against a torch fp32 brute-force
            scan of the ORIGINAL fp32 rows (regenerated from the seeds; test infrastructure, outside the timed region)
  verified  sortedness + planted neighbours on every query, the parity block, every sweep entry equal to the headline's
            answers, e2e equal to the device path; ``verified: false`` makes the exit code 1
  cpu_baseline / --impl reference
            The reference's own path is chromadb's HNSW (a third-party wheel that is not in this image), so
            the CPU arm is the oracle port of the exact scan the north star names as the reference: numpy fp32
            ``Q @ C.T`` + top-k with all BLAS threads, and the OpenMP C restatement; the faster of the two is reported.
            It is timed on a bounded row sample and scaled linearly to the full corpus (an exact scan is linear in rows).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

DIM = 384
CHUNK_ROWS = 500_000          # global generation chunk; seed = collection seed + chunk index
MMA_GROUP = 256               # queries per K2 corpus pass above 128 (CTA pairs)
PARITY_QUERIES = 64
METRIC = "QPS exact top-10 over 100Mx384 bf16 (cosine), row-sharded"

# BASELINE.json configs.  cfg4 is the headline (its metric string is the one BASELINE.json quotes); QPS is quoted at the top
# of its batch range, the other batch sizes are measured briefly in the same run ("sweep").
CONFIGS = {
    "cfg4": {"rows": 100_000_000, "batch": 4096, "k": 10, "sweep": "1,1s,8,64,128,256,1024", "seeds": [1234],
             "what": "synthetic unit-norm {rows}x384 {dtype} corpus, cosine top-{k}, query batch {batch}"},
    "cfg2": {"rows": 10_000_000, "batch": 1024, "k": 10, "sweep": "1,64", "seeds": [1234],
             "what": "synthetic {rows}x384 {dtype} corpus, cosine top-{k}, query batch {batch} (sweep: 1, 64)"},
    "cfg3": {"rows": 10_000_000, "batch": 1024, "k": 10, "k_each": 50, "sweep": "1,64", "seeds": [1234, 2234],
             "what": "dual-encoder ensemble, 2 x {rows}x384 {dtype}, per-collection top-50 -> RRF(60) -> top-{k}, "
                     "query batch {batch} (sweep: 1, 64)"},
    "cfg5": {"rows_per_gpu": 125_000_000, "batch": 1024, "k": 100, "sweep": "", "seeds": [1234],
             "what": "synthetic {rows}x384 {dtype} corpus sharded over the GPUs (125M rows = 96 GB each), cosine top-{k}, "
                     "query batch {batch}"},
}


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="cfg4", choices=sorted(CONFIGS))
    ap.add_argument("--rows", type=int, default=0, help="0 = the config's")
    ap.add_argument("--batch", type=int, default=0)
    ap.add_argument("--k", type=int, default=0)
    ap.add_argument("--dtype", default="bf16", choices=["bf16", "f32"])
    ap.add_argument("--path", default="auto", choices=["auto", "stream", "mma"])
    ap.add_argument("--sweep", default=None, help="extra batch sizes measured briefly (comma list, '' = none)")
    ap.add_argument("--single-process", action="store_true",
                    help="one process owns all --gpus devices (the store API's mode) instead of one process per GPU")
    ap.add_argument("--exchange", default="auto", choices=["auto", "nccl", "copy", "peer"])
    ap.add_argument("--cpu-sample-rows", type=int, default=0, help="0 = sized for a few seconds of CPU work per step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-parity", action="store_true", help="skip the fp32 reference scan of the parity block")
    args = ap.parse_args()
    cfg = CONFIGS[args.config]
    if not args.rows:
        args.rows = cfg.get("rows") or cfg["rows_per_gpu"] * max(1, args.gpus)
    args.batch = args.batch or cfg["batch"]
    args.k = args.k or cfg["k"]
    if args.sweep is None:
        args.sweep = cfg["sweep"]
    args.seeds = cfg["seeds"]
    args.k_each = cfg.get("k_each", args.k)
    return args


# ---------------------------------------------------------------------------------------------
def metric_name(args):
    if args.config == "cfg4":
        return METRIC
    return f"QPS exact top-{args.k} ({args.config}), row-sharded"


def workload_name(args):
    return f"{args.config}: " + CONFIGS[args.config]["what"].format(rows=args.rows, dtype=args.dtype, k=args.k, batch=args.batch)


def gen_chunk(torch, device, seed, c, rows):
    g = torch.Generator(device=device).manual_seed(seed + c)
    return torch.randn((rows, DIM), generator=g, device=device, dtype=torch.float32)


def planted_rows(n_rows, batch):
    import numpy as np

    return np.random.default_rng(99).integers(0, n_rows, size=batch)


def make_queries(torch, device, n_rows, batch, seed=1234, qseed=4321):
    """Gaussian queries; every even one is a noisy copy of a corpus row."""
    g = torch.Generator(device=device).manual_seed(qseed)
    q = torch.randn((batch, DIM), generator=g, device=device, dtype=torch.float32)
    noise = torch.randn((batch, DIM), generator=g, device=device, dtype=torch.float32)
    rows = planted_rows(n_rows, batch)
    by_chunk = {}
    for i in range(0, batch, 2):
        by_chunk.setdefault(int(rows[i]) // CHUNK_ROWS, []).append(i)
    for c, idxs in by_chunk.items():
        lo = c * CHUNK_ROWS
        chunk = gen_chunk(torch, device, seed, c, min(CHUNK_ROWS, n_rows - lo))
        for i in idxs:
            q[i] = chunk[int(rows[i]) - lo] + 0.1 * noise[i]
        del chunk
    return q, rows


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, device_index):
        self.samples = []
        self.proc = None
        self.dev = device_index

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits", "-lms", "100",
                 "-i", str(self.dev)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            p = [x.strip() for x in line.split(",")]
            if len(p) >= 7:
                self.samples.append(p)

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.12)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        mhz, mx, reasons = [], None, set()
        for p in self.samples:
            try:
                mhz.append(float(p[0]))
                mx = float(p[1])
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), p[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(mhz) if mhz else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(mhz)}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(p) as f:
            d = json.load(f)
        return (float(d["hbm_gbs"]), float(d.get("bf16_tflops_sustained") or d["bf16_tflops"]),
                float(d["bf16_tflops"]), "measured")
    except Exception:
        return 6650.0, 1400.0, 1590.0, "fallback"


def ncu_traffic(path_kind, rows_local):
    """DRAM bytes per scan launch (dram__bytes_read.sum + dram__bytes_write.sum from the committed
    ``ncu --set full`` capture under profiles/, measured per row), scaled to this launch; or None."""
    try:
        with open(os.path.join(ROOT, "profiles", "roofline_traffic.json")) as f:
            return float(json.load(f)[path_kind]["bytes_per_row"]) * rows_local
    except Exception:
        return None


def scan_kernel_of(batch, path, k, dtype="bf16", rows_local=None):
    """Which scan kernel a batch runs on (mirrors search_on_stream in csrc/api.cu) and the queries
    one launch of it serves."""
    mma_ok = k <= 100  # fp32 collections select on a bf16 copy of their rows
    use_mma = mma_ok and path in ("mma", "auto")
    if use_mma and path == "auto" and rows_local is not None:  # small shards are launch-bound: K1 is 3 launches
        if (batch == 1 and rows_local <= 2_000_000) or (batch <= 4 and rows_local <= 200_000):
            use_mma = False
    if not use_mma:
        return "scan_stream_kernel", "stream", min(batch, 4)
    if batch <= 64 and (k <= 64 or batch <= 32):  # (64 queries x 256 candidates do not fit K2s's shared-memory lists)
        nq = 16 if batch <= 16 else (32 if batch <= 32 else 64)
        return f"scan_mma_small_kernel<{nq},*> (tcgen05, corpus rows as M, queries as N)", "mma_small", batch
    if batch <= 128:
        return "scan_mma_kernel<1,1> (tcgen05, one CTA per SM)", "mma_cg1", batch
    return "scan_mma_kernel<1,2> (tcgen05 cta_group::2, CTA pairs)", "mma_cg2", None  # per launch: from the count


def roofline_of(batch, path, k, rows_local, elem, scan_ms, scan_launches, searches, step_ms_total):
    """Roofline of the dominant (scan) kernel from its CUDA-event time inside the library.
    Algorithmic work per launch: bytes = rows_per_gpu * 384 * sizeof(elem) -- the corpus is
    read once per launch whatever the number of queries; flops = 2 * rows_per_gpu * 384 * queries the
    launch serves.  The bound is whichever of bytes/hbm_peak and flops/tensor_peak is the longer."""
    dtype = "bf16" if elem == 2 else "f32"
    hbm_peak, tf_sustained, tf_burst, peak_kind = measured_peaks()
    kernel, kind, _ = scan_kernel_of(batch, path, k, dtype, rows_local)
    avg_launch_s = (scan_ms / max(scan_launches, 1)) / 1e3
    q_per_launch = batch * max(searches, 1) / max(scan_launches, 1)
    bytes_per_launch = rows_local * DIM * elem
    flops_per_launch = 2.0 * rows_local * DIM * q_per_launch
    gbs = bytes_per_launch / avg_launch_s / 1e9
    tfs = flops_per_launch / avg_launch_s / 1e12
    t_hbm, t_tensor = bytes_per_launch / (hbm_peak * 1e9), flops_per_launch / (tf_sustained * 1e12)
    tensor_bound = kind != "stream" and t_tensor > t_hbm
    return {
        "bound": "tensor" if tensor_bound else "hbm",
        "achieved": tfs if tensor_bound else gbs,
        "peak": tf_sustained if tensor_bound else hbm_peak,
        "unit": "TFLOP/s" if tensor_bound else "GB/s",
        "frac": (tfs / tf_sustained) if tensor_bound else (gbs / hbm_peak),
        "traffic": ncu_traffic(kind + (f"_co{int(round(q_per_launch / MMA_GROUP))}"
                                       if kind == "mma_cg2" and q_per_launch > MMA_GROUP else ""), rows_local),
        "peak_kind": (f"{peak_kind} (MEASURED_PEAKS.json bf16_tflops_sustained: the kernel runs inside a long step; "
                      f"burst {tf_burst})" if tensor_bound else f"{peak_kind} (MEASURED_PEAKS.json hbm_gbs)"),
        "kernel": kernel,
        "queries_per_launch": q_per_launch,
        "bytes_per_launch": bytes_per_launch,
        "flops_per_launch": flops_per_launch,
        "avg_launch_ms": avg_launch_s * 1e3,
        "launches_timed": scan_launches,
        "hbm_gbs": gbs, "hbm_frac": gbs / hbm_peak,
        "tensor_tflops": tfs, "tensor_frac_sustained": tfs / tf_sustained, "tensor_frac_burst": tfs / tf_burst,
        "scan_share_of_step": scan_ms / step_ms_total if step_ms_total else None,
    }


# ---------------------------------------------------------------------------------------------
def cpu_scan_qps(n_rows_full, batch, k, sample_rows, steps, warmup, torch=None, device=None):
    """Exact fp32 scan on the host cores over a bounded row sample; QPS scaled to the full corpus."""
    import numpy as np

    from oracle import cscan
    from oracle import exact_scan as ox

    if not sample_rows:
        # ~1.5e12 flop (sgemm regime) or 2M rows (bandwidth regime) per step: a few seconds on a host CPU
        sample_rows = int(max(250_000, min(2_000_000, 1.5e12 / (2.0 * DIM * batch))))
        sample_rows = (sample_rows // 250_000) * 250_000
    sample_rows = int(min(sample_rows, n_rows_full))
    rng_rows = sample_rows
    # the same synthetic rows as the GPU arm where a device is available, else numpy Gaussians
    if torch is not None and device is not None:
        parts = []
        for c in range((rng_rows + CHUNK_ROWS - 1) // CHUNK_ROWS):
            parts.append(gen_chunk(torch, device, 1234, c, min(CHUNK_ROWS, rng_rows - c * CHUNK_ROWS)).cpu().numpy())
        corpus = np.concatenate(parts)
        q = make_queries(torch, device, sample_rows, batch)[0].cpu().numpy()
    else:
        rng = np.random.default_rng(1234)
        corpus = rng.standard_normal((rng_rows, DIM), dtype=np.float32)
        q = rng.standard_normal((batch, DIM), dtype=np.float32)
    corpus = ox.prepare_corpus(corpus, "cosine", "f32")
    q = ox.prepare_queries(q, "cosine")
    cores = os.cpu_count() or 1
    # torchrun exports OMP_NUM_THREADS=1 to every rank; the CPU arm is meant to use all the host threads it can
    try:
        from threadpoolctl import threadpool_limits

        threadpool_limits(limits=cores)
    except Exception as e:  # noqa: BLE001
        sys.stderr.write(f"[bench] threadpoolctl unavailable ({e}); BLAS keeps its default thread count\n")

    def run_numpy():
        return ox.exact_topk_thresholded(q, corpus, k, "cosine", chunk_rows=max(16384, min(1 << 20, (1 << 26) // batch)))

    def run_c():
        return cscan.exact_topk_prepared(corpus, q, k, "cosine", nthreads=cores)

    def run_numpy_mt():  # queries dealt to one worker thread per core in slices, BLAS single-threaded inside each
        try:
            from threadpoolctl import threadpool_limits as tl

            with tl(limits=1):
                return ox.exact_topk_thresholded_mt(q, corpus, k, "cosine", threads=cores)
        except ImportError:
            return ox.exact_topk_thresholded_mt(q, corpus, k, "cosine", threads=cores)

    results = {}
    # large batches: the sliced scan (one big sgemm leaves the element-wise passes on one core, 3-4 x slower);
    # small ones: one sgemm with all BLAS threads, and the C port, which scans row by row for each query
    arms = [("numpy_sgemm_sliced", run_numpy_mt)] if batch >= 256 else [("numpy_sgemm", run_numpy)]
    if batch <= 16:
        arms.append(("c_openmp", run_c))
    for name, fn in arms:
        try:
            for _ in range(max(1, min(warmup, 2))):
                fn()
            ts = []
            for _ in range(max(1, steps)):
                t0 = time.perf_counter()
                fn()
                ts.append(time.perf_counter() - t0)
            results[name] = statistics.median(ts)
        except Exception as e:  # the C port needs gcc/OpenMP on the box
            results[name] = None
            sys.stderr.write(f"[bench] cpu arm {name} unavailable: {e}\n")
    best = min((t, n) for n, t in results.items() if t)
    t_full = best[0] * (n_rows_full / sample_rows)
    return {
        "value": batch / t_full,
        "unit": "queries/s",
        "cores": cores,
        "kind": "port",
        "sample": (f"exact fp32 scan of {sample_rows} of {n_rows_full} rows x {batch} queries, "
                   f"{best[1]} ({', '.join(f'{n}={t * 1e3:.0f}ms' for n, t in results.items() if t)}), "
                   f"time scaled linearly to the full corpus"),
        "ms_per_step_sample": best[0] * 1e3,
        "sample_rows": sample_rows,
    }


def run_reference(args, out=sys.stdout):
    """--impl reference: the CPU exact scan (oracle port; chromadb is not installable offline)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    torch = device = None
    try:
        import torch as _t

        if _t.cuda.is_available():
            torch, device = _t, _t.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
    except Exception:
        pass
    # bounded: a few seconds per step at the default sample
    steps = max(1, min(args.steps, 8))
    total_rows = args.rows * len(args.seeds)  # the ensemble scans every collection
    r = cpu_scan_qps(total_rows, args.batch, args.k_each, args.cpu_sample_rows, steps, args.warmup, torch, device)
    line = {
        "impl": "reference",
        "metric": metric_name(args), "value": r["value"], "unit": "queries/s", "n_gpus": args.gpus,
        "steps": steps, "warmup": args.warmup,
        "ms_per_step": r["ms_per_step_sample"] * (total_rows / r["sample_rows"]),
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(args), "rows": args.rows, "batch": args.batch, "k": args.k,
                   "note": "reference arm = CPU exact scan (oracle port); Chroma HNSW is not runnable offline"},
        "cpu_baseline": {k_: r[k_] for k_ in ("value", "unit", "cores", "kind", "sample")},
        "e2e": {"value": r["value"], "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), file=out, flush=True)


# ---------------------------------------------------------------------------------------------
class Collection:
    """One synthetic collection on this process's GPUs: the plain fr_index on one GPU, else the row-sharded fr_group
    (one process per GPU under torchrun, or one process owning all of them with --single-process)."""

    def __init__(self, frb, torch, dist, args, world, rank, local_devices, seed):
        self.torch, self.frb, self.seed, self.n = torch, frb, seed, args.rows
        self.local_devices = local_devices
        self.shards_total = world if world > 1 else len(local_devices)
        self.first = rank if world > 1 else 0
        per_shard = (self.n + self.shards_total - 1) // self.shards_total
        if self.shards_total == 1:
            self.ix = frb.ShardIndex(dim=DIM, space="cosine", dtype=args.dtype, device=local_devices[0], reserve_rows=self.n)
            self.grp = None
            self.shards = [self.ix]
        else:
            if world > 1:
                self.grp = frb.ShardGroup.from_torch_distributed(dim=DIM, space="cosine", dtype=args.dtype,
                                                                 device=local_devices[0], reserve_rows=self.n)
            else:
                self.grp = frb.ShardGroup(dim=DIM, space="cosine", dtype=args.dtype, devices=local_devices,
                                          reserve_rows=self.n, exchange=args.exchange)
            self.ix = None
            self.shards = [self.grp.shard(j) for j in range(len(local_devices))]
        self.rows_local = per_shard
        for s in self.shards:
            s.set_path(args.path)
            if os.environ.get("FR_MMA_CO_GROUPS"):  # tuning experiments only
                s.set_option("mma_co_groups", int(os.environ["FR_MMA_CO_GROUPS"]))

    def load(self):
        """Cyclic placement of the global rows: row r -> shard r % W (keys = global rows)."""
        torch, n, w = self.torch, self.n, self.shards_total
        t0 = time.time()
        for c in range((n + CHUNK_ROWS - 1) // CHUNK_ROWS):
            r0 = c * CHUNK_ROWS
            rows = min(CHUNK_ROWS, n - r0)
            for j, d in enumerate(self.local_devices):
                device = torch.device("cuda", d)
                chunk = gen_chunk(torch, device, self.seed, c, rows)
                if w == 1:
                    self.shards[j].append_device(chunk, None, first_key=r0)
                else:
                    off = (self.first + j - r0) % w
                    keys = torch.arange(r0 + off, r0 + rows, w, device=device, dtype=torch.int64)
                    self.shards[j].append_device(chunk[off::w].contiguous(), keys)
                del chunk
        for d in self.local_devices:
            torch.cuda.synchronize(d)
        if self.grp is not None:
            self.grp.adopt_rows(n)
        self.rows_local = max(s.rows() for s in self.shards)
        return time.time() - t0

    def search_device(self, qs, k, out_d=None, out_k=None):
        """qs: one query block per local device.  Merged result on the first local device."""
        if self.grp is None:
            return self.ix.search_device(qs[0], k, out_d, out_k)
        n = len(self.local_devices)
        d, kk = self.grp.search_device(qs, k, [out_d] + [None] * (n - 1), [out_k] + [None] * (n - 1), merge_on=[0])
        return d[0], kk[0]

    def search_host(self, q_ptr, b, k, d_ptr, k_ptr):
        (self.ix or self.grp).search_raw(q_ptr, b, k, d_ptr, k_ptr)

    def profile(self, on):
        for s in self.shards:
            if on:
                s.profile_read()
            s.set_profile(on)

    def profile_read(self):
        """(scan ms, launches, searches) of the slowest local shard."""
        reads = [s.profile_read() for s in self.shards]
        return max(reads, key=lambda r: r[0])

    def stat(self, name):
        return sum(s.stat(name) for s in self.shards)

    def close(self):
        (self.ix or self.grp).close()


def fp32_reference(torch, dist, device, rank, world, n, seed, q_raw, k, got_keys):
    """Test infrastructure: the north star's reference -- an fp32 brute-force scan of the ORIGINAL fp32 rows (regenerated
    from the seeds, normalised like the reference: x / (|x| + 1e-30)), in plain torch, chunks dealt round-robin to the
    ranks.  Returns on every rank (numpy): ref_keys, ref_scores [Bq, k]; the fp32 scores of the ids the GPU path returned;
    and sum_i |q_i c_i| of both (the bf16 storage bound is 2^-9 times that)."""
    torch.backends.cuda.matmul.allow_tf32 = False
    q = q_raw / (q_raw.norm(dim=1, keepdim=True) + 1e-30)
    qa = q.abs()
    bq = q.shape[0]
    best_s = torch.full((bq, k), float("-inf"), device=device)
    best_k = torch.full((bq, k), -1, dtype=torch.int64, device=device)
    best_a = torch.zeros((bq, k), device=device)
    gk = torch.from_numpy(got_keys).to(device)
    got_s = torch.zeros((bq, k), device=device)
    got_a = torch.zeros((bq, k), device=device)
    for c in range(rank, (n + CHUNK_ROWS - 1) // CHUNK_ROWS, world):
        base = c * CHUNK_ROWS
        rows = min(CHUNK_ROWS, n - base)
        x = gen_chunk(torch, device, seed, c, rows)
        x = x / (x.norm(dim=1, keepdim=True) + 1e-30)
        s = q @ x.T
        a = qa @ x.abs().T
        m = (gk >= base) & (gk < base + rows)
        idx = (gk - base).clamp(0, rows - 1)
        got_s += torch.where(m, s.gather(1, idx), torch.zeros_like(got_s))
        got_a += torch.where(m, a.gather(1, idx), torch.zeros_like(got_a))
        ts, ti = s.topk(min(k, rows), dim=1)
        cs, ck = torch.cat([best_s, ts], 1), torch.cat([best_k, ti + base], 1)
        ca = torch.cat([best_a, a.gather(1, ti)], 1)
        o = cs.argsort(dim=1, descending=True, stable=True)[:, :k]
        best_s, best_k, best_a = cs.gather(1, o), ck.gather(1, o), ca.gather(1, o)
        del x, s, a
    if world > 1:
        parts = [[torch.empty_like(t) for _ in range(world)] for t in (best_s, best_k, best_a)]
        for p, t in zip(parts, (best_s, best_k, best_a)):
            dist.all_gather(p, t)
        cs, ck, ca = (torch.cat(p, 1) for p in parts)
        o = cs.argsort(dim=1, descending=True, stable=True)[:, :k]
        best_s, best_k, best_a = cs.gather(1, o), ck.gather(1, o), ca.gather(1, o)
        dist.all_reduce(got_s)
        dist.all_reduce(got_a)
    return tuple(t.cpu().numpy() for t in (best_k, best_s, best_a, got_s, got_a))


def north_star_gate(got_keys, got_dist, ref, storage, rows):
    """The north star's gate, literally (ids bit-exact except where the two candidates' fp32 scores tie within 1e-3
    relative; scores within 1e-3 relative for bf16 storage, 1e-5 for fp32) and with the rigorous bf16 storage bound
    2^-9 * sum|q_i c_i| added as absolute slack (any correct bf16 index needs it on data whose top scores are ~0.3)."""
    import numpy as np

    ref_k, ref_s, ref_a, got_s, got_a = ref
    rtol = 1e-3 if storage == "bf16" else 1e-5
    diff_id = got_keys != ref_k
    gap = np.abs(got_s - ref_s)
    scale = np.maximum(np.abs(got_s), np.abs(ref_s))
    slack = (2.0 ** -9) * (got_a + ref_a) if storage == "bf16" else 0.0
    lit_id = diff_id & (gap > 1e-3 * scale + 1e-12)
    slk_id = diff_id & (gap > 1e-3 * scale + slack + 1e-12)
    score = 1.0 - got_dist.astype(np.float64)
    err = np.abs(score - got_s)
    rel = err / np.maximum(np.abs(got_s), 1e-30)
    lit_sc = err > rtol * np.maximum(np.abs(score), np.abs(got_s)) + 2e-6
    slk_sc = err > rtol * np.maximum(np.abs(score), np.abs(got_s)) + 2e-6 + ((2.0 ** -9) * got_a if storage == "bf16" else 0.0)
    return {
        "rows": rows, "queries": int(got_keys.shape[0]), "k": int(got_keys.shape[1]),
        "reference": "torch fp32 brute-force scan of the original fp32 rows (regenerated from the seeds)",
        "id_mismatches": int(diff_id.sum()),
        "id_mismatches_outside_1e-3_ties": int(lit_id.sum()),
        "id_mismatches_outside_ties_and_bf16_storage_bound": int(slk_id.sum()),
        "max_rel_score_err": float(rel.max()),
        "max_abs_score_err": float(err.max()),
        "scores_outside_1e-3_rel" if storage == "bf16" else "scores_outside_1e-5_rel": int(lit_sc.sum()),
        "scores_outside_rel_and_bf16_storage_bound": int(slk_sc.sum()),
        "pass": bool(slk_id.sum() == 0 and slk_sc.sum() == 0),
        "pass_literal": bool(lit_id.sum() == 0 and lit_sc.sum() == 0),
    }


def rrf_check(np, key_lists, fused_keys, fused_scores, k_rrf, k_out):
    """Checker for the fused output (retriever.py:94-107 restated inline; bench.py may not use oracle/ here)."""
    for b in range(fused_keys.shape[0]):
        agg = {}
        for lst in key_lists:
            for r, key in enumerate(lst[b].tolist()):
                if key != -1:
                    agg[key] = agg.get(key, 0.0) + 1.0 / (k_rrf + r + 1)
        want = sorted(agg.items(), key=lambda it: it[1], reverse=True)[:k_out]
        got = [(int(a), float(s)) for a, s in zip(fused_keys[b], fused_scores[b]) if a != -1]
        if got != want:
            return False
    return True


def preflight(frb, torch, dist, world, rank, local_devices, out_info):
    """N > 1: before anything is timed, the sharded store path must give the one-GPU answer on this box -- a small
    collection ingested through the SPMD / multi-device group vs a plain one-GPU index on this rank's first device."""
    import numpy as np

    rng = np.random.default_rng(2024)
    n = 60_000
    corpus = rng.standard_normal((n, DIM), dtype=np.float32)
    corpus[n - 1] = corpus[0]
    keys = np.arange(n, dtype=np.int64) + 11
    if world > 1:
        grp = frb.ShardGroup.from_torch_distributed(dim=DIM, dtype="bf16", device=local_devices[0])
    else:
        grp = frb.ShardGroup(dim=DIM, dtype="bf16", devices=local_devices)
    one = frb.ShardIndex(dim=DIM, dtype="bf16", device=local_devices[0])
    for lo in range(0, n, 20_000):
        grp.upsert(corpus[lo:lo + 20_000], keys[lo:lo + 20_000])
        one.upsert(corpus[lo:lo + 20_000], keys[lo:lo + 20_000])
    ok = grp.count() == n
    for b, k in ((1, 10), (64, 10), (300, 10), (130, 100)):
        q = rng.standard_normal((b, DIM), dtype=np.float32)
        q[0] = corpus[0]
        wd, wk = one.search(q, k)
        gd, gk = np.empty_like(wd), np.empty_like(wk)
        qq = np.ascontiguousarray(q)
        grp.search_raw(qq.ctypes.data if (rank == 0 or world == 1) else None, b, k, gd.ctypes.data, gk.ctypes.data)
        ok &= bool((gk == wk).all()) and float(np.abs(gd - wd).max()) <= 2e-6
        ok &= int(gk[0, 0]) == 11 and int(gk[0, 1]) == 11 + n - 1  # the tie comes back in insertion order
    out_info["exchange"] = grp.exchange
    out_info["nccl_version"] = grp.info("nccl_version")
    grp.close()
    one.close()
    if world > 1:
        flag = torch.tensor([1 if ok else 0], device=torch.device("cuda", local_devices[0]))
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        ok = bool(flag.item())
    return ok


# ---------------------------------------------------------------------------------------------
def run_ours(args, out=sys.stdout):
    import numpy as np
    import torch
    import torch.distributed as dist

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the B200 backend has no CPU fallback")
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world == 1 and args.gpus > 1 and not args.single_process:
        raise SystemExit(f"--gpus {args.gpus} needs torchrun with --nproc-per-node {args.gpus} (or --single-process)")
    if world > 1:
        local_devices = [local_rank]
    else:
        local_devices = list(range(args.gpus)) if args.single_process else [0]
    n_gpus = world if world > 1 else len(local_devices)
    torch.cuda.set_device(local_devices[0])
    device = torch.device("cuda", local_devices[0])
    devs = [torch.device("cuda", d) for d in local_devices]
    if world > 1:
        dist.init_process_group("nccl", device_id=device)

    import financial_rag_b200 as frb
    from financial_rag_b200 import _lib

    n, b, k = args.rows, args.batch, args.k
    elem = 2 if args.dtype == "bf16" else 4

    def barrier():
        if world > 1:
            dist.barrier()
        for d in local_devices:
            torch.cuda.synchronize(d)

    # ---- pre-flight: the sharded path answers like one GPU (N > 1) -------------------------------------------------
    pre = {"ran": False}
    if n_gpus > 1:
        pre = {"ran": True}
        pre["ok"] = preflight(frb, torch, dist, world, rank, local_devices, pre)

    # ---- build the collection(s) -----------------------------------------------------------------------------------
    cols = [Collection(frb, torch, dist, args, world, rank, local_devices, seed) for seed in args.seeds]
    t_build = sum(c.load() for c in cols)
    rows_local = cols[0].rows_local
    ensemble = len(cols) > 1
    k_each = args.k_each if ensemble else k
    if os.environ.get("FR_BENCH_CORRUPT_SHARD"):  # negative control: the run must fail its verification
        victims = torch.tensor([int(r) for r in planted_rows(n, 64)[0:64:2]], dtype=torch.int64).numpy()
        for s in cols[0].shards:
            s.delete(victims)

    sweep_items = [x.strip() for x in args.sweep.split(",") if x.strip()]
    max_b = max([b] + [int(x.rstrip("s")) for x in sweep_items])
    # one query set, sliced per batch size: every entry answers (a prefix of) the same queries
    q_all, planted_all = [], None
    for ci, seed in enumerate(args.seeds):
        q0, pl = make_queries(torch, device, n, max_b, seed=seed, qseed=4321 + ci)
        q_all.append([q0 if d == device else q0.to(d) for d in devs])
        planted_all = pl if planted_all is None else planted_all
    fuse = frb.EnsembleSearcher([c.ix or c.grp for c in cols], k_each=k_each, k_rrf=60, k_out=k) if ensemble else None

    def search_dev(batch):
        if ensemble:
            return fuse.search_device([(qs[0][:batch] if c.grp is None else [t[:batch] for t in qs])
                                       for c, qs in zip(cols, q_all)])
        return cols[0].search_device([t[:batch] for t in q_all[0]], k)

    def timed(fn, steps):
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in devs]
        for (e0, _), d in zip(ev, local_devices):
            e0.record(torch.cuda.current_stream(d))
        for _ in range(steps):
            r = fn()
        for (_, e1), d in zip(ev, local_devices):
            e1.record(torch.cuda.current_stream(d))
        for d in local_devices:
            torch.cuda.synchronize(d)
        ms = torch.tensor([max(e0.elapsed_time(e1) for e0, e1 in ev)], device=device)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item()), r

    def measure(batch, steps, warmup, profile=False):
        for _ in range(warmup):
            search_dev(batch)
        barrier()
        if profile:
            for c in cols:
                c.profile(True)
        l0 = _lib.launch_count()
        ms, (d, kk) = timed(lambda: search_dev(batch), steps)
        launches = _lib.launch_count() - l0
        scan = None
        if profile:
            for c in cols:
                c.profile(False)
            reads = [c.profile_read() for c in cols]
            scan = tuple(sum(r[i] for r in reads) for i in range(3))
        barrier()
        return ms, launches, scan, (d.clone(), kk.clone())

    # pinned host buffers of the e2e legs (allocated once, for the largest batch)
    q_pins = []
    for qs in q_all:
        t = torch.empty((max_b, DIM), dtype=torch.float32).pin_memory()
        t.copy_(qs[0].cpu())
        q_pins.append(t)
    od_pins = [torch.empty((max_b, k), dtype=torch.float32).pin_memory()]
    ok_pins = [torch.empty((max_b, k), dtype=torch.int64).pin_memory()]
    feeds = rank == 0

    def e2e_once(batch):
        """Host buffers through the public call: H2D, [broadcast,] scan, exchange, merge, [fusion,] D2H."""
        if ensemble:  # EnsembleSearcher.search: pinned query blocks in, fused [B, k] out; the lists never leave the GPUs
            return fuse.search([qp[:batch] for qp in q_pins])
        cols[0].search_host(q_pins[0].data_ptr() if feeds else None, batch, k, od_pins[0].data_ptr(), ok_pins[0].data_ptr())
        return None

    def measure_e2e(batch, steps):
        for _ in range(3):
            e2e_once(batch)
        barrier()
        ms, r = timed(lambda: e2e_once(batch), steps)
        barrier()
        return ms, r

    # ---- headline: device-resident queries -------------------------------------------------------------------------
    sampler = ClockSampler(local_devices[0])
    if rank == 0:
        sampler.start()
    ms, launches, scan, (d, kk) = measure(b, args.steps, max(args.warmup, 3), profile=True)
    clocks = sampler.stop() if rank == 0 else None
    value = b * args.steps / (ms / 1e3)

    # ---- verification of what was timed (outside the timed region) -------------------------------------------------
    kk_h, d_h = kk.cpu().numpy(), d.cpu().numpy()
    checks = {}
    planted = planted_all[:b]
    if ensemble:
        checks["fused_scores_descending"] = bool((np.diff(d_h, axis=1) <= 0).all())
        # both encoders hold the same planted rows: the fused winner of an even query is its planted row, from both lists
        checks["planted_neighbours_first"] = all(int(kk_h[i, 0]) == int(planted[i]) and d_h[i, 0] == 2.0 / 61
                                                 for i in range(0, b, 2))
        lists = []
        for c, qs in zip(cols, q_all):
            ld, lk = c.search_device([t[:min(b, 64)] for t in qs], k_each)
            lists.append(lk.cpu().numpy())
        checks["fusion_equals_python_rrf"] = rrf_check(np, lists, kk_h[:min(b, 64)], d_h[:min(b, 64)], 60, k)
    else:
        checks["distances_ascending"] = bool((np.diff(d_h, axis=1) >= 0).all())
        checks["planted_neighbours_first"] = all(int(kk_h[i, 0]) == int(planted[i]) and (1.0 - float(d_h[i, 0])) > 0.9
                                                 for i in range(0, b, 2))
        checks["random_queries_score_low"] = all((1.0 - float(d_h[i, 0])) < 0.6 for i in range(1, b, 2))
    parity = None
    if not args.no_parity:
        bq = min(b, PARITY_QUERIES)
        if ensemble:  # gate each collection's top-k' lists (fewer queries: k' = 50 columns each)
            parity = []
            for c, qs, seed in zip(cols, q_all, args.seeds):
                ld, lk = c.search_device([t[:min(bq, 16)] for t in qs], k_each)
                ld, lk = ld.cpu().numpy(), lk.cpu().numpy()
                ref = fp32_reference(torch, dist, device, rank, world, n, seed, qs[0][:min(bq, 16)], k_each, lk)
                parity.append(north_star_gate(lk, ld, ref, args.dtype, n))
            checks["north_star_gate"] = all(p["pass"] for p in parity)
        else:
            ref = fp32_reference(torch, dist, device, rank, world, n, args.seeds[0], q_all[0][0][:bq], k, kk_h[:bq])
            parity = north_star_gate(kk_h[:bq], d_h[:bq], ref, args.dtype, n)
            checks["north_star_gate"] = parity["pass"]

    # ---- e2e: host buffers through the public call -----------------------------------------------------------------
    ms_e2e, fused_host = measure_e2e(b, args.steps)
    e2e_value = b * args.steps / (ms_e2e / 1e3)
    if ensemble:
        checks["e2e_equals_device_path"] = bool((fused_host[1] == kk_h).all())
    else:
        checks["e2e_equals_device_path"] = bool((ok_pins[0].numpy()[:b] == kk_h).all())

    # ---- roofline of the scan kernel -------------------------------------------------------------------------------
    scan_ms, scan_launches, scan_searches = scan
    roofline = roofline_of(b, args.path, k_each, rows_local, elem, scan_ms, scan_launches, scan_searches, ms)
    uncertified = sum(c.stat("mma_uncertified_queries") for c in cols)

    # ---- brief sweep over the other batch sizes (each with its own roofline, e2e and answer check) -----------------
    sweep = []
    for item in sweep_items:
        spath = "stream" if item.endswith("s") else args.path
        sb = int(item.rstrip("s"))
        if sb == b and spath == args.path:
            continue
        for c in cols:
            for s in c.shards:
                s.set_path(spath)
        # size each entry to ~0.5 s of device time (>= 3 steps) after a short pause, so that it is neither a cold
        # burst nor riding on the power state the previous (much heavier or lighter) entry left behind
        probe_ms, _, _, _ = measure(sb, 2, 3)
        st = max(3, min(60, int(500.0 / max(probe_ms / 2, 1e-3))))
        time.sleep(0.5)
        sms, _, sscan, (sd, sk) = measure(sb, st, 3, profile=True)
        e_st = max(3, st // 2)
        ems, _ = measure_e2e(sb, e_st)
        for c in cols:
            for s in c.shards:
                s.set_path(args.path)
        r = roofline_of(sb, spath, k_each, rows_local, elem, sscan[0], sscan[1], sscan[2], sms)
        m = min(sb, b)
        same = bool((sk.cpu().numpy()[:m] == kk_h[:m]).all()) and \
            float(np.abs(sd.cpu().numpy()[:m].astype(np.float64) - d_h[:m]).max()) <= 2e-6
        sweep.append({"batch": sb, "path": spath, "steps": st, "qps": sb * st / (sms / 1e3), "ms_per_step": sms / st,
                      "e2e_qps": sb * e_st / (ems / 1e3), "e2e_ms_per_step": ems / e_st,
                      "answers_equal_headline": same,
                      "kernel": r["kernel"], "bound": r["bound"], "frac": r["frac"], "achieved": r["achieved"],
                      "unit": r["unit"], "hbm_gbs": r["hbm_gbs"], "hbm_frac": r["hbm_frac"],
                      "tensor_tflops": r["tensor_tflops"], "tensor_frac_sustained": r["tensor_frac_sustained"],
                      "scan_share_of_step": r["scan_share_of_step"]})
    if sweep:
        checks["sweep_answers_equal_headline"] = all(e["answers_equal_headline"] for e in sweep)
    if pre["ran"]:
        checks["preflight_sharded_equals_one_gpu"] = pre["ok"]
    verified = all(checks.values())

    # ---- CPU baseline (rank 0, N = 1 only) -------------------------------------------------------------------------
    cpu = None
    if rank == 0 and n_gpus == 1 and not args.no_cpu_baseline:
        r = cpu_scan_qps(n * len(cols), b, k_each, args.cpu_sample_rows, 5, 1, torch, device)
        cpu = {k_: r[k_] for k_ in ("value", "unit", "cores", "kind", "sample")}

    if rank == 0:
        par = "row-shard x%d (%s)" % (n_gpus, "one GPU" if n_gpus == 1 else
                                      ("one process per GPU, fr_group over NCCL" if world > 1 else
                                       f"one process, fr_group exchange={cols[0].grp.exchange}"))
        line = {
            "metric": metric_name(args), "value": value, "unit": "queries/s", "n_gpus": n_gpus, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": args.dtype, "data": "synthetic",
            "config": {"workload": workload_name(args), "rows": n, "rows_per_gpu": rows_local, "dim": DIM,
                       "collections": len(cols), "batch": b, "k": k, "k_per_collection": k_each, "space": "cosine",
                       "parallelism": par, "placement": "cyclic (row r on shard r % N)", "path": args.path,
                       "l2_policy": "inputs larger than L2 (>= 9.6 GB per GPU), no flush",
                       "build_s": round(t_build, 2), "verified": verified, "nccl_version": pre.get("nccl_version")},
            "e2e": {"value": e2e_value, "unit": "queries/s", "h2d_bytes_per_step": b * DIM * 4 * len(cols),
                    "d2h_bytes_per_step": b * k * (16 if ensemble else 12),
                    "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": launches,
            "roofline": roofline,
            "cpu_baseline": cpu,
            "clocks": clocks,
            "verified": verified,
            "checks": checks,
            "parity": parity,
            "mma_uncertified_queries": uncertified,
            "mma_rescanned_queries": sum(c.stat("mma_rescanned_queries") for c in cols),
            "reference_chroma": {"recall_at_10": None, "qps": None,
                                 "note": "Chroma HNSW CPU path not runnable offline: the chromadb wheel is neither in the "
                                         "reference tree nor in this image (SURVEY.md 8c/8d); the CPU arm is the exact scan"},
            "sweep": sweep,
        }
        print(json.dumps(line), file=out, flush=True)
        if not verified:
            sys.stderr.write(f"[bench] VERIFICATION FAILED: {json.dumps(checks)}\n")
    for c in cols:
        c.close()
    if world > 1:
        dist.destroy_process_group()
    return 0 if verified else 1


def main():
    args = parse_args()
    # stdout must carry exactly ONE JSON line, but libraries print there too (NCCL_DEBUG=INFO logs, NCCL's version
    # banner): park fd 1 on stderr for the whole run and hand the real stdout only to the final print.
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    out = os.fdopen(real_stdout, "w")
    rc = 0
    try:
        if args.impl == "reference":
            run_reference(args, out)
        else:
            rc = run_ours(args, out)
    finally:
        sys.stdout.flush()
        out.flush()
    sys.exit(rc)


if __name__ == "__main__":
    main()
