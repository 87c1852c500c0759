// C ABI of libfrb200.so (include/fr_index.h): host-side shard management around the kernels.
//
// One fr_index = one collection shard resident on one B200:
//     corpus  [cap_rows][dim]  bf16 or fp32, row-major, rows in insertion order (+32 rows of pad
//                              so the streaming kernel may read whole 32-row blocks)
//     keys    [cap_rows]       int64, INT64_MIN marks a deleted row
// plus grow-only scratch (query staging, per-CTA partial lists, result staging, pinned mirror).
// Mirrors what ChromaChildStore asks of chromadb (parent_child/chroma_child_store.py:32-80).
#include <atomic>
#include <condition_variable>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <new>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/fr_index.h"
#include "fr_common.cuh"
#include "fr_host.h"
#include "fr_kernels.h"

#include <execinfo.h>
#include <signal.h>
#include <unistd.h>

namespace {
// FRB200_SEGV_TRACE=1: print the native backtrace of a crashing thread (python's faulthandler only sees the ctypes call)
void segv_trace(int sig) {
    void *frames[64];
    const int n = backtrace(frames, 64);
    const char msg[] = "\n[libfrb200] fatal signal, native backtrace:\n";
    if (write(2, msg, sizeof(msg) - 1) < 0) {}
    backtrace_symbols_fd(frames, n, 2);
    signal(sig, SIG_DFL);
    raise(sig);
}
struct SegvTraceInstall {
    SegvTraceInstall() {
        const char *e = getenv("FRB200_SEGV_TRACE");
        if (e && e[0] == '1') {
            signal(SIGSEGV, segv_trace);
            signal(SIGBUS, segv_trace);
            signal(SIGABRT, segv_trace);
        }
    }
} g_segv_trace_install;
thread_local std::string g_last_error;
std::atomic<int64_t> g_launches{0};
std::mutex g_dev_mu;
fr::DevInfo g_dev[64];
}  // namespace

namespace fr {

int fail(int code, const char *fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_last_error = buf;
    return code;
}

std::shared_mutex &graph_wait_mutex() {
    static std::shared_mutex m;
    return m;
}

int check_device(int device, int *sm_count) {
    std::lock_guard<std::mutex> lk(g_dev_mu);
    if (device >= 0 && device < 64 && g_dev[device].state == 1) {
        if (sm_count) *sm_count = g_dev[device].sm_count;
        return FR_OK;
    }
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        cudaGetLastError();
        return fail(FR_ENODEV, "no CUDA device available (%s); this backend has no CPU fallback",
                    e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
    }
    if (device < 0 || device >= n || device >= 64) return fail(FR_EINVAL, "device %d out of range (0..%d)", device, n - 1);
    DevInfo &d = g_dev[device];
    if (d.state == 0) {
        int major = 0, minor = 0, sms = 0;
        e = cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, device);
        if (e == cudaSuccess) e = cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, device);
        if (e == cudaSuccess) e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
        if (e != cudaSuccess) return fail(FR_ECUDA, "cudaDeviceGetAttribute: %s", cudaGetErrorString(e));
        d.major = major;
        d.minor = minor;
        d.sm_count = sms;
        d.state = (major == 10) ? 1 : -1;
    }
    if (d.state != 1)
        return fail(FR_ENODEV, "device %d is sm_%d%d; libfrb200 is built for sm_100a (B200) only", device, d.major,
                    d.minor);
    if (sm_count) *sm_count = d.sm_count;
    return FR_OK;
}

}  // namespace fr

namespace {

using fr::check_device;
using fr::DevBuf;
using fr::DeviceGuard;
using fr::fail;
using fr::PinBuf;

constexpr int64_t PAD_ROWS = 32;

}  // namespace

namespace fr {
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }
void count_launches(int64_t n) { g_launches.fetch_add(n, std::memory_order_relaxed); }
}  // namespace fr

struct fr_index {
    int dim = 0, metric = FR_COSINE, dtype = FR_BF16, device = 0, sm_count = 148;
    int path = FR_PATH_AUTO;
    int64_t rows = 0;      // physical rows, deleted ones included
    int64_t cap_rows = 0;  // usable rows (allocation holds cap_rows + PAD_ROWS)
    int64_t n_deleted = 0;
    uint8_t *corpus = nullptr;
    int64_t *keys = nullptr;
    // fp32 collections (cosine, width 384): bf16 copy of the rows that the tensor-core scans SELECT from (the exact
    // scores always come from the fp32 rows).  Built lazily on the first batched search, extended after appends,
    // rebuilt after an in-place overwrite.
    uint8_t *shadow = nullptr;
    int64_t shadow_cap = 0, shadow_rows = 0;
    bool shadow_dirty = false;
    int mma_f32_shadow = 1;
    cudaStream_t stream = nullptr;   // host-path stream
    cudaEvent_t last_use = nullptr;  // orders scratch reuse across caller streams
    DevBuf q_raw, q_prep, q_keys, partials, out_dist, out_keys, stage_vecs, stage_keys, stage_rows;
    DevBuf q_bf16, err_bound, sel, sel_keys, flags, fail, fb_partials, tau, hist, progress, s_lists;  // K2 path
    DevBuf cmax;             // inner-product / l2 collections: largest row norm (device float), the scale of the error bounds
    int64_t cmax_rows = 0;   // rows it covers; an in-place overwrite resets it
    DevBuf norm2;            // l2 collections: |c|^2 per row (the tensor-core selection ranks by 2 q.c - |c|^2)
    int64_t norm2_cap = 0;
    DevBuf kth_exact, r_q, r_misc, r_tau, r_partials, r_sel, r_sel_keys;     // K2 second-chance pass
    int mma_min_batch = 2;  // FR_PATH_AUTO: batches at least this large go to the tensor-core scans; smaller ones
                            // too when the swapped-operand kernel K2s serves them (it out-streams K1: TMA ring)
    int mma_small_max = 64;  // K2s serves batches up to this size (0 = never)
    int mma_split = -1;      // K2s reads the queries as two bf16 terms: 1 always, 0 never, -1 up to mma_split_max queries
    int mma_split_max = 32;  // (measured: free up to 32 queries -- one MMA of N = 2 x 32 per K step; 64 do not fit an accumulator)
    int64_t small_rows_b1 = 2000000, small_rows_b4 = 200000;  // FR_PATH_AUTO: below these sizes batch 1 / batch <= 4 take K1
    int mma_bound_scale_pct = 100;  // diagnostics / tests: certification error bounds x this / 100 (>= 100: stricter, still exact)
    int mma_score_hist = 1; // K2 / K2s: share a score histogram between the CTAs (0 = threshold slots only; for A/B timing)
    int mma_debug = 0;      // diagnostics (scripts/ablate_mma.py): results are wrong when non-zero
    int mma_wide_lists = 1; // k in (64, 100]: keep 256 candidates per query instead of 128
    int mma_max_lead = 6;   // K2 co-resident groups: tiles a group may run ahead of the slowest group of its stream (0 = unthrottled)
    int mma_co_groups = 8;  // K2: at most this many query groups of 256 share one corpus stream through L2, kept together
                            // by the lead throttle (100M rows: batch 1024 = 4 groups, one HBM pass per 1024 queries,
                            // 14.4k -> 17.0k QPS; batch 4096 with 8 groups another +2 % over 4: the board is power-bound
                            // and every HBM byte not fetched is clock for the tensor cores)
    int mma_retry_blocks = 0;  // second-chance blocks enqueued per call: 0 = as many as the last finished search needed (at
                               // least one; a block = 3 launches that exit at once when nobody failed into it -- a 4096-query
                               // search used to enqueue 32 of them, ~100 empty launches), -1 = one per 128 queries, n = fixed
    int retry_blocks_cur = 1;
    int *fail_mirror = nullptr;  // pinned: first-pass failure count of the last finished search (written by retry_prep_kernel)
    DevBuf stats;           // [0] queries K2 could not certify (re-scanned by the stream kernel), cumulative
    int64_t n_searches = 0, n_queries = 0, n_mma_queries = 0;
    // Host entry point fr_index_search: concurrent callers (Flask threads, api_server.py:1366-1371) hold `mu` only while
    // their call is ENQUEUED on the shard's stream; each then waits for its own event and copies its results out of its own
    // pinned slot, so the host copies and waits of one search overlap the GPU work of the next.
    struct HostSlot {
        PinBuf pin;
        cudaEvent_t done = nullptr;
        bool busy = false;
    };
    static constexpr int N_SLOTS = 4;
    HostSlot slots[N_SLOTS];
    std::mutex slot_mu;
    std::condition_variable slot_cv;
    std::mutex mu;
    bool profile = false;
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> prof_events;  // one pair per profiled search
    std::vector<cudaEvent_t> prof_pool;
    int64_t prof_launches = 0;
    // host-path search graphs (small collections: the search is a chain of ~3-12 short kernels whose launch gaps,
    // not their run time, set the latency): one captured graph per (B, k), replayed while nothing it baked in changed
    struct SearchGraph {
        cudaGraphExec_t exec = nullptr;
        uint64_t state = 0;      // state_hash() the graph was captured under
        uint64_t seen = 0;       // state_hash() after the last eager run of this shape (capture on the second stable call)
        int64_t launches = 0, d_searches = 0, d_queries = 0, d_mma_queries = 0;
        bool failed = false;     // capture did not work for this shape: stay eager
    };
    std::unordered_map<uint64_t, SearchGraph> graphs;
    int use_graphs = 1;
    int host_debug = 0;      // diagnostics (scripts/stress_concurrent.py): 1 = thread-local capture mode, 2 = the shard lock is held through the wait, 4 = graph work and waits may overlap (crashes)
    int64_t graph_max_bytes = int64_t(2) << 30;  // corpora above this are bandwidth-bound: launch gaps do not matter
    int64_t n_graph_replays = 0;
    std::unordered_map<int64_t, int64_t> keymap;  // key -> row (live rows only)
    bool keymap_valid = true;

    size_t row_bytes() const { return static_cast<size_t>(dim) * (dtype == FR_BF16 ? 2 : 4); }
};

namespace {

int grow(fr_index *ix, int64_t need_rows, bool exact) {
    if (need_rows <= ix->cap_rows) return FR_OK;
    int64_t new_cap = need_rows;
    if (!exact) {
        const int64_t geo = ix->cap_rows + ix->cap_rows / 2;
        if (geo > new_cap) new_cap = geo;
        if (new_cap < 1024) new_cap = 1024;
    }
    uint8_t *nc = nullptr;
    int64_t *nk = nullptr;
    const size_t rb = ix->row_bytes();
    cudaError_t e = cudaMalloc(reinterpret_cast<void **>(&nc), static_cast<size_t>(new_cap + PAD_ROWS) * rb);
    if (e != cudaSuccess && new_cap != need_rows) {
        cudaGetLastError();
        new_cap = need_rows;
        e = cudaMalloc(reinterpret_cast<void **>(&nc), static_cast<size_t>(new_cap + PAD_ROWS) * rb);
    }
    if (e != cudaSuccess) {
        cudaGetLastError();
        return fail(FR_ENOMEM, "cannot allocate %lld rows x %zu B on device %d: %s", (long long)new_cap, rb,
                    ix->device, cudaGetErrorString(e));
    }
    e = cudaMalloc(reinterpret_cast<void **>(&nk), static_cast<size_t>(new_cap + PAD_ROWS) * sizeof(int64_t));
    if (e != cudaSuccess) {
        cudaGetLastError();
        cudaFree(nc);
        return fail(FR_ENOMEM, "cannot allocate key array for %lld rows: %s", (long long)new_cap, cudaGetErrorString(e));
    }
    // all earlier work on this shard must be finished before the old arrays go away
    FR_CUDA(cudaDeviceSynchronize());
    FR_CUDA(cudaMemsetAsync(nc + static_cast<size_t>(ix->rows) * rb, 0,
                            static_cast<size_t>(new_cap + PAD_ROWS - ix->rows) * rb, ix->stream));
    if (ix->rows > 0) {
        FR_CUDA(cudaMemcpyAsync(nc, ix->corpus, static_cast<size_t>(ix->rows) * rb, cudaMemcpyDeviceToDevice, ix->stream));
        FR_CUDA(cudaMemcpyAsync(nk, ix->keys, static_cast<size_t>(ix->rows) * sizeof(int64_t), cudaMemcpyDeviceToDevice,
                                ix->stream));
    }
    FR_CUDA(cudaStreamSynchronize(ix->stream));
    if (ix->corpus) cudaFree(ix->corpus);
    if (ix->keys) cudaFree(ix->keys);
    if (ix->shadow) cudaFree(ix->shadow);  // sized for the old capacity: rebuilt on the next batched search
    ix->shadow = nullptr;
    ix->shadow_cap = ix->shadow_rows = 0;
    ix->corpus = nc;
    ix->keys = nk;
    ix->cap_rows = new_cap;
    return FR_OK;
}

// stream `s` is about to touch the shard's scratch: wait for whoever used it last
int begin_use(fr_index *ix, cudaStream_t s) {
    FR_CUDA(cudaStreamWaitEvent(s, ix->last_use, 0));
    return FR_OK;
}
int end_use(fr_index *ix, cudaStream_t s) {
    FR_CUDA(cudaEventRecord(ix->last_use, s));
    return FR_OK;
}

int rebuild_keymap(fr_index *ix) {
    if (ix->keymap_valid) return FR_OK;
    std::vector<int64_t> hk(static_cast<size_t>(ix->rows));
    FR_CUDA(cudaDeviceSynchronize());
    if (ix->rows > 0)
        FR_CUDA(cudaMemcpy(hk.data(), ix->keys, hk.size() * sizeof(int64_t), cudaMemcpyDeviceToHost));
    ix->keymap.clear();
    ix->keymap.reserve(hk.size() * 2);
    for (int64_t r = 0; r < ix->rows; ++r)
        if (hk[r] != fr::KEY_TOMBSTONE) ix->keymap[hk[r]] = r;
    ix->keymap_valid = true;
    return FR_OK;
}

// profiling brackets around the scan launches (bench.py's roofline line)
struct ProfScope {
    fr_index *ix;
    cudaStream_t s;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    int64_t launches_before = 0;
    int begin() {
        if (!ix->profile) return FR_OK;
        for (cudaEvent_t *ev : {&ev0, &ev1}) {
            if (!ix->prof_pool.empty()) {
                *ev = ix->prof_pool.back();
                ix->prof_pool.pop_back();
            } else {
                FR_CUDA(cudaEventCreate(ev));
            }
        }
        FR_CUDA(cudaEventRecord(ev0, s));
        launches_before = fr_launch_count();
        return FR_OK;
    }
    int end() {
        if (!ix->profile) return FR_OK;
        FR_CUDA(cudaEventRecord(ev1, s));
        ix->prof_events.emplace_back(ev0, ev1);
        ix->prof_launches += fr_launch_count() - launches_before;
        return FR_OK;
    }
};

// The tensor-core scans serve bf16 cosine collections: K2 and K2s at width 384, K2s alone (small batches, larger
// ones in slices) at width 768 -- the multi-vector store's bert-base token vectors (multivector_store.py:70).
// K2s reads the queries as ONE bf16 term (selection error up to |q - bf16(q)|, ~1e-3) or as TWO (hi + lo, error
// ~1e-5: nearly every query certified at once even where thousands of rows score within 1e-2 of the best).
bool shadow_serves(const fr_index *ix) {
    return ix->dtype == FR_F32 && ix->mma_f32_shadow && (ix->metric == FR_COSINE || ix->metric == FR_IP) && ix->dim == 384;
}

// inner-product collections: bring the largest row norm up to date (stream-ordered on `s`)
int ensure_cmax(fr_index *ix, cudaStream_t s) {
    if (ix->metric != FR_IP && ix->metric != FR_L2) return FR_OK;
    if (!ix->cmax.p) {
        FR_CUDA(ix->cmax.need(64));
        ix->cmax_rows = 0;
    }
    const bool want_norm2 = ix->metric == FR_L2;
    if (want_norm2 && ix->norm2_cap < ix->cap_rows + PAD_ROWS) {  // the shard grew: recompute into a larger array
        FR_CUDA(cudaStreamSynchronize(s));
        ix->norm2.release();
        FR_CUDA(ix->norm2.need(static_cast<size_t>(ix->cap_rows + PAD_ROWS) * sizeof(float)));
        ix->norm2_cap = ix->cap_rows + PAD_ROWS;
        ix->cmax_rows = 0;
    }
    if (ix->cmax_rows == 0) FR_CUDA(cudaMemsetAsync(ix->cmax.p, 0, 64, s));
    if (ix->cmax_rows < ix->rows) {
        const size_t rb = ix->row_bytes();
        FR_CUDA(fr::launch_row_norm_max(ix->corpus + static_cast<size_t>(ix->cmax_rows) * rb, ix->dtype == FR_BF16,
                                        ix->rows - ix->cmax_rows, ix->dim, static_cast<float *>(ix->cmax.p),
                                        want_norm2 ? static_cast<float *>(ix->norm2.p) + ix->cmax_rows : nullptr, s));
        ix->cmax_rows = ix->rows;
    }
    return FR_OK;
}

// bring the bf16 selection copy of an fp32 collection up to date (stream-ordered on `s`)
int ensure_shadow(fr_index *ix, cudaStream_t s) {
    const size_t rb16 = static_cast<size_t>(ix->dim) * 2;
    if (ix->shadow == nullptr || ix->shadow_cap < ix->cap_rows) {
        if (ix->shadow) {
            FR_CUDA(cudaStreamSynchronize(s));
            cudaFree(ix->shadow);
            ix->shadow = nullptr;
        }
        cudaError_t e = cudaMalloc(reinterpret_cast<void **>(&ix->shadow), static_cast<size_t>(ix->cap_rows + PAD_ROWS) * rb16);
        if (e != cudaSuccess) {
            cudaGetLastError();
            ix->shadow_cap = ix->shadow_rows = 0;
            return fail(FR_ENOMEM, "cannot allocate the bf16 selection copy of %lld fp32 rows: %s", (long long)ix->cap_rows,
                        cudaGetErrorString(e));
        }
        ix->shadow_cap = ix->cap_rows;
        ix->shadow_rows = 0;
        FR_CUDA(cudaMemsetAsync(ix->shadow, 0, static_cast<size_t>(ix->cap_rows + PAD_ROWS) * rb16, s));
    }
    if (ix->shadow_dirty) {
        ix->shadow_rows = 0;
        ix->shadow_dirty = false;
    }
    if (ix->shadow_rows < ix->rows) {
        const int64_t first = ix->shadow_rows;
        FR_CUDA(fr::launch_shadow_convert(reinterpret_cast<const float *>(ix->corpus) + first * ix->dim,
                                          ix->shadow + static_cast<size_t>(first) * rb16, (ix->rows - first) * ix->dim, s));
        ix->shadow_rows = ix->rows;
    }
    return FR_OK;
}

int small_split(const fr_index *ix, int B, int ksel) {
    // two terms while they fit one MMA per K step (measured: no cost at all, profiles/r01_sweep_split_10m.jsonl);
    // always at widths without a second-chance pass (certify in the first)
    const bool want = ix->dim != 384 || ix->mma_split == 1 || (ix->mma_split < 0 && B <= ix->mma_split_max);
    return (want && fr::scan_mma_small_nq(B, ksel, ix->dim, 1) != 0) ? 1 : 0;
}
bool small_serves(const fr_index *ix, int B, int ksel) {
    if (ix->mma_small_max <= 0 || B > ix->mma_small_max) return false;
    return fr::scan_mma_small_nq(B, ksel, ix->dim, small_split(ix, B, ksel)) != 0;
}
int mma_slice(const fr_index *ix, int k) {  // 0 = not eligible, else the largest batch one pass may take
    const int ksel = fr::scan_mma_ksel(k, ix->mma_wide_lists);
    if (ksel == 0 || ix->rows <= 0) return 0;
    if (ix->metric == FR_L2) {  // l2: the swapped-operand kernel only (its epilogue subtracts the row norms), in slices it holds
        if (ix->dtype != FR_BF16 || ix->dim != 384) return 0;
        const int m = fr::scan_mma_small_max_batch(ksel, ix->dim, 0);
        return m < ix->mma_small_max ? m : ix->mma_small_max;
    }
    if (ix->dtype != FR_BF16) return shadow_serves(ix) ? 1 << 30 : 0;  // fp32 rows: selection on a bf16 copy
    if (ix->dim == 384) return 1 << 30;
    if (ix->dim != 768) return 0;  // the re-scan safety net exists for 384 and 768 only
    const int m = fr::scan_mma_small_max_batch(ksel, ix->dim, 1);
    return m < ix->mma_small_max ? m : ix->mma_small_max;
}
bool mma_eligible(const fr_index *ix, int k) { return mma_slice(ix, k) > 0; }

// K1 path: CUDA-core streaming scan, ceil(B/4) corpus passes.
int search_stream(fr_index *ix, const float *q, int B, int k, float *d_out_dist, uint64_t *d_out_packed,
                  int64_t *d_out_keys, cudaStream_t s) {
    fr::ScanArgs sa{};
    sa.corpus = ix->corpus;
    sa.keys_or_null = ix->n_deleted > 0 ? ix->keys : nullptr;
    sa.queries = q;
    sa.n_rows = ix->rows;
    sa.dim = ix->dim;
    sa.bf16 = ix->dtype == FR_BF16;
    sa.l2 = ix->metric == FR_L2;
    sa.k = k;
    sa.nq_total = B;
    sa.stream = s;
    sa.grid = fr::scan_stream_plan_grid(sa, ix->sm_count);
    FR_CUDA(ix->partials.need(static_cast<size_t>(sa.grid) * B * k * sizeof(uint64_t)));
    sa.partials = static_cast<uint64_t *>(ix->partials.p);
    ProfScope prof{ix, s};
    int rc = prof.begin();
    if (rc != FR_OK) return rc;
    FR_CUDA(fr::launch_scan_stream(sa));
    rc = prof.end();
    if (rc != FR_OK) return rc;

    fr::MergeArgs ma{};
    ma.packed = sa.partials;
    ma.P = sa.grid;
    ma.shard_stride = static_cast<int64_t>(B) * k;
    ma.B = B;
    ma.k = k;
    ma.shards = false;
    ma.row_keys = ix->keys;
    ma.l2 = sa.l2;
    ma.out_dist = d_out_dist;
    ma.out_packed = d_out_packed;
    ma.out_keys = d_out_keys;
    ma.stream = s;
    FR_CUDA(fr::launch_merge_topk(ma));
    return FR_OK;
}

// Second-chance blocks the next search enqueues (see fr_index::mma_retry_blocks).  A pure function of the option and of
// the pinned failure count, so calling it twice for one search changes nothing.
int plan_retry_blocks(fr_index *ix) {
    const int R = fr::scan_mma_retry_max();
    if (ix->mma_retry_blocks != 0) {
        ix->retry_blocks_cur = ix->mma_retry_blocks;
    } else {
        const int last = ix->fail_mirror ? *static_cast<volatile int *>(ix->fail_mirror) : 0;
        const int need = (last + R - 1) / R;
        ix->retry_blocks_cur = need < 1 ? 1 : need;
    }
    return ix->retry_blocks_cur;
}

// K2 path: tensor-core selection of k' candidates per query, exact fp32-query rescoring with
// certification, a second tensor-core pass for what could not be certified, and a device-side re-scan
// of whatever is still open after that.
int search_mma(fr_index *ix, const float *d_raw_queries, int B, int k, float *d_out_dist, uint64_t *d_out_packed,
               int64_t *d_out_keys, cudaStream_t s) {
    const int ksel = fr::scan_mma_ksel(k, ix->mma_wide_lists);
    const int group = fr::scan_mma_group(B);
    const int nq_pad = ((B + group - 1) / group) * group;
    // small batches take the swapped-operand kernel (tensor work proportional to the batch)
    const bool small = small_serves(ix, B, ksel);
    const int split = small ? small_split(ix, B, ksel) : 0;
    if (!small && ix->dim != 384) return fail(FR_EUNSUP, "internal: width %d needs the small-batch kernel", ix->dim);
    const bool l2 = ix->metric == FR_L2;
    if (l2 && !small) return fail(FR_EUNSUP, "internal: l2 collections take the small-batch kernel only");
    const bool second_chance = ix->dim == 384 && !l2;  // the second-chance pass runs on K2 (384-wide, dot products)
    const fr::MmaPlan plan = fr::scan_mma_plan(ix->sm_count, ix->rows, B, ix->mma_co_groups);
    const int grid = plan.lists_max;  // partial lists per query (at most)
    // second-chance blocks: R queries each, enough of them for every query of the call
    const int R = fr::scan_mma_retry_max();
    int slices = 0;
    if (second_chance) {
        const int all = (B + R - 1) / R, want = plan_retry_blocks(ix);
        slices = (want < 0 || want > all) ? all : want;
        if (!ix->fail_mirror) {
            FR_CUDA(cudaHostAlloc(reinterpret_cast<void **>(&ix->fail_mirror), 64, cudaHostAllocPortable));
            ix->fail_mirror[0] = 0;
        }
    }
    FR_CUDA(ix->q_bf16.need(static_cast<size_t>(nq_pad) * ix->dim * 2 * (1 + split)));
    FR_CUDA(ix->err_bound.need(6 * static_cast<size_t>(B) * sizeof(float)));  // |e| one-term | two-term | |e.q| one-term | two-term | 1/scale | |q|^2
    FR_CUDA(ix->partials.need(static_cast<size_t>(grid) * B * ksel * sizeof(uint64_t)));
    FR_CUDA(ix->sel.need(static_cast<size_t>(B) * ksel * sizeof(uint64_t)));
    FR_CUDA(ix->sel_keys.need(static_cast<size_t>(B) * ksel * sizeof(int64_t)));
    FR_CUDA(ix->flags.need(static_cast<size_t>(B)));
    // 2 counters (+2 pad) | fail_list [B] | fail_list2 [B] | retry_n [slices]
    FR_CUDA(ix->fail.need((4 + 2 * static_cast<size_t>(B) + static_cast<size_t>(slices)) * sizeof(int)));
    FR_CUDA(ix->tau.need(static_cast<size_t>(B) * ksel * sizeof(uint32_t)));
    // cosine collections: the CTAs also share a score histogram per query (scan_mma_small.cu, scan_mma.cu)
    const bool use_hist = ix->metric == FR_COSINE && ix->mma_score_hist != 0;
    if (use_hist) FR_CUDA(ix->hist.need(static_cast<size_t>(B) * fr::SCORE_HIST_WORDS * sizeof(uint32_t)));
    FR_CUDA(ix->q_prep.need(static_cast<size_t>(B) * ix->dim * sizeof(float)));
    if (!ix->stats.p) {
        FR_CUDA(ix->stats.need(64));
        FR_CUDA(cudaMemsetAsync(ix->stats.p, 0, 64, s));
    }
    ix->n_mma_queries += B;
    int *counters = static_cast<int *>(ix->fail.p);
    int *fail_count = counters, *fail_count2 = counters + 1;
    int *fail_list = counters + 4, *fail_list2 = fail_list + B, *retry_n = fail_list2 + B;
    float *eb_one = static_cast<float *>(ix->err_bound.p), *eb_two = eb_one + B, *ea_one = eb_two + B, *ea_two = ea_one + B;
    // one launch: cosine normalisation (K0's arithmetic), bf16 copy, error bounds, and the reset of this call's
    // threshold slots and failure counters
    fr::PrepArgs pa{};
    pa.raw = d_raw_queries;
    pa.nq = B;
    pa.nq_pad = nq_pad;
    pa.dim = ix->dim;
    pa.q_prep = static_cast<float *>(ix->q_prep.p);
    pa.qb = ix->q_bf16.p;
    pa.err_bound = eb_one;
    pa.err_bound_split = eb_two;
    pa.err_alpha = ea_one;
    pa.err_alpha_split = ea_two;
    pa.split = split;
    pa.bound_scale = static_cast<float>(ix->mma_bound_scale_pct) / 100.0f;
    pa.tau_g = static_cast<uint32_t *>(ix->tau.p);
    pa.ksel = (ix->mma_debug & 128) ? 0 : ksel;  // diagnostics: 128 = the threshold slots keep the previous search's values
    pa.hist = use_hist ? static_cast<uint32_t *>(ix->hist.p) : nullptr;
    pa.counters = counters;
    pa.n_counters = 2;
    pa.normalize = ix->metric == FR_COSINE;
    float *inv_scale = eb_one + 4 * static_cast<size_t>(B);
    if (ix->metric != FR_COSINE) {
        int rcm = ensure_cmax(ix, s);
        if (rcm != FR_OK) return rcm;
        pa.cmax = static_cast<const float *>(ix->cmax.p);
    }
    pa.inv_scale = inv_scale;
    pa.qnorm2 = inv_scale + B;
    pa.stream = s;
    FR_CUDA(fr::launch_prep_queries(pa));
    const float *q = pa.q_prep;

    const bool f32_rows = ix->dtype == FR_F32;  // search_on_stream has brought the bf16 copy up to date
    if (f32_rows && (ix->shadow == nullptr || ix->shadow_rows != ix->rows || ix->shadow_dirty))
        return fail(FR_ECUDA, "internal: the bf16 selection copy is not up to date");
    // |q . (c - bf16(c))| <= 2^-9 sum|q_i c_i| <= 2^-9 |q||c| for an fp32 unit row c
    const float extra_bound = f32_rows ? 0.001953125f * 1.001f : 0.0f;
    fr::MmaScanArgs ms{};
    ms.corpus = f32_rows ? ix->shadow : ix->corpus;
    ms.dim = ix->dim;
    ms.keys_or_null = ix->n_deleted > 0 ? ix->keys : nullptr;
    ms.queries_bf16 = ix->q_bf16.p;
    ms.split = split;
    ms.nq_pad = nq_pad;
    ms.n_rows = ix->rows;
    ms.nq_total = B;
    ms.ksel = ksel;
    ms.partials = static_cast<uint64_t *>(ix->partials.p);
    ms.plan = plan;
    ms.dbg = ix->mma_debug;
    if (!small && plan.co > 1 && ix->mma_max_lead > 0) {
        FR_CUDA(ix->progress.need(static_cast<size_t>(plan.lists_max) * plan.co * sizeof(uint32_t)));
        ms.progress = static_cast<uint32_t *>(ix->progress.p);
        ms.max_lead = ix->mma_max_lead;
    }
    ms.tau_g = static_cast<uint32_t *>(ix->tau.p);
    ms.hist = pa.hist;
    ms.norm2 = l2 ? static_cast<const float *>(ix->norm2.p) : nullptr;
    ms.stream = s;
    ProfScope prof{ix, s};
    int rc = prof.begin();
    if (rc != FR_OK) return rc;
    if (small)
        FR_CUDA(fr::launch_scan_mma_small(ms, 0, B));
    else
        FR_CUDA(fr::launch_scan_mma(ms));
    rc = prof.end();
    if (rc != FR_OK) return rc;

    // queries of the full launches hold plan.lists partial lists each, those of the tail launch plan.lists_tail
    for (int part = 0; part < 2; ++part) {
        const int q0 = part == 0 ? 0 : plan.tail_q0;
        const int nq = part == 0 ? plan.tail_q0 : B - plan.tail_q0;
        if (nq <= 0) continue;
        fr::MergeArgs ma{};
        ma.packed = ms.partials + static_cast<size_t>(q0) * ksel;
        ma.P = part == 0 ? plan.lists : plan.lists_tail;
        ma.shard_stride = static_cast<int64_t>(B) * ksel;
        ma.B = nq;
        ma.k = ksel;
        ma.shards = false;
        ma.row_keys = ix->keys;
        ma.l2 = false;  // (selection lists carry packed keys only; distances come out of the rescore pass)
        ma.out_packed = static_cast<uint64_t *>(ix->sel.p) + static_cast<size_t>(q0) * ksel;
        ma.out_keys = static_cast<int64_t *>(ix->sel_keys.p) + static_cast<size_t>(q0) * ksel;
        ma.stream = s;
        FR_CUDA(fr::launch_merge_topk(ma));
    }

    unsigned long long *stat_uncertified = static_cast<unsigned long long *>(ix->stats.p);
    unsigned long long *stat_rescanned = stat_uncertified + 1;

    // first rescore pass: exact fp32-query scores of the k' candidates, certification
    fr::RescoreArgs ra{};
    ra.sel = static_cast<const uint64_t *>(ix->sel.p);
    ra.ksel = ksel;
    ra.queries = q;
    ra.dim = ix->dim;
    ra.corpus = ix->corpus;
    ra.f32_rows = f32_rows ? 1 : 0;
    ra.extra_bound = extra_bound;
    ra.row_keys = ix->keys;
    ra.err_bound = split ? eb_two : eb_one;
    ra.err_alpha = split ? ea_two : ea_one;
    ra.inv_scale = inv_scale;
    ra.l2 = l2 ? 1 : 0;
    ra.qnorm2 = pa.qnorm2;
    ra.cmax = pa.cmax;
    ra.split = split;
    ra.B = B;
    ra.k = k;
    ra.out_dist = d_out_dist;
    ra.out_packed = d_out_packed;
    ra.out_keys = d_out_keys;
    ra.flags = static_cast<uint8_t *>(ix->flags.p);
    ra.fail_count = second_chance ? fail_count : fail_count2;  // without a second chance: straight to the re-scan
    ra.fail_list = second_chance ? fail_list : fail_list2;
    ra.fail_total = stat_uncertified;
    ra.fail_total2 = second_chance ? nullptr : stat_rescanned;
    FR_CUDA(ix->kth_exact.need(static_cast<size_t>(B) * sizeof(float)));
    ra.kth_exact = static_cast<float *>(ix->kth_exact.p);
    ra.stream = s;
    FR_CUDA(fr::launch_rescore(ra));

    if (second_chance) {
        // second chance on the tensor cores for what could not be certified: gather the failures into blocks of
        // R queries, then per block scan above a fixed threshold -> merge -> rescore.  Everything is enqueued
        // unconditionally; the launches of a block nobody failed into return at once (no host round trip).
        const int ksel_r = fr::scan_mma_retry_ksel(ksel);
        const fr::MmaPlan rplan = fr::scan_mma_plan(ix->sm_count, ix->rows, R, 1);
        const size_t slots = static_cast<size_t>(slices) * R;
        FR_CUDA(ix->r_q.need(slots * ix->dim * 2));
        FR_CUDA(ix->r_misc.need(slots * sizeof(float)));
        FR_CUDA(ix->r_tau.need(slots * ksel_r * sizeof(uint32_t)));
        FR_CUDA(ix->r_partials.need(static_cast<size_t>(rplan.lists_max) * R * ksel_r * sizeof(uint64_t)));  // reused by every block
        FR_CUDA(ix->r_sel.need(static_cast<size_t>(R) * ksel_r * sizeof(uint64_t)));
        FR_CUDA(ix->r_sel_keys.need(static_cast<size_t>(R) * ksel_r * sizeof(int64_t)));
        float *tau0 = static_cast<float *>(ix->r_misc.p);

        fr::RetryPrepArgs rp{};
        rp.queries = q;
        rp.err_bound = eb_one;  // the second-chance scan reads one-term bf16 queries
        rp.err_alpha = ea_one;
        rp.extra_bound = extra_bound;
        rp.inv_scale = inv_scale;
        rp.kth_exact = ra.kth_exact;
        rp.fail_count = fail_count;
        rp.fail_list = fail_list;
        rp.slices = slices;
        rp.fail_count2 = fail_count2;
        rp.fail_list2 = fail_list2;
        rp.rescanned_total = stat_rescanned;
        rp.host_mirror = ix->fail_mirror;
        rp.qb_retry = ix->r_q.p;
        rp.tau0 = tau0;
        rp.retry_n = retry_n;
        rp.tau_g_retry = static_cast<uint32_t *>(ix->r_tau.p);
        rp.ksel = ksel_r;
        rp.flags = ra.flags;
        rp.stream = s;
        FR_CUDA(fr::launch_retry_prep(rp));

        for (int sl = 0; sl < slices; ++sl) {
            fr::MmaScanArgs rs = ms;
            rs.queries_bf16 = static_cast<const uint8_t *>(ix->r_q.p) + static_cast<size_t>(sl) * R * ix->dim * 2;
            rs.split = 0;
            rs.nq_pad = R;
            rs.nq_total = R;
            rs.ksel = ksel_r;
            rs.partials = static_cast<uint64_t *>(ix->r_partials.p);
            rs.plan = rplan;
            rs.tau_g = rp.tau_g_retry + static_cast<size_t>(sl) * ksel_r * R;
            rs.hist = nullptr;
            rs.nq_dev = retry_n + sl;
            rs.tau0 = tau0 + static_cast<size_t>(sl) * R;
            FR_CUDA(fr::launch_scan_mma(rs));

            fr::MergeArgs mr{};
            mr.packed = rs.partials;
            mr.P = rplan.lists;
            mr.shard_stride = static_cast<int64_t>(R) * ksel_r;
            mr.B = R;
            mr.k = ksel_r;
            mr.shards = false;
            mr.row_keys = ix->keys;
            mr.l2 = false;
            mr.out_packed = static_cast<uint64_t *>(ix->r_sel.p);
            mr.out_keys = static_cast<int64_t *>(ix->r_sel_keys.p);
            mr.limit = retry_n + sl;
            mr.stream = s;
            FR_CUDA(fr::launch_merge_topk(mr));

            fr::RescoreArgs rr = ra;
            rr.sel = static_cast<const uint64_t *>(ix->r_sel.p);
            rr.ksel = ksel_r;
            rr.err_bound = eb_one;
            rr.err_alpha = ea_one;
            rr.split = 0;
            rr.B = R;
            rr.fail_count = fail_count2;
            rr.fail_list = fail_list2;
            rr.fail_total = stat_rescanned;
            rr.fail_total2 = nullptr;
            rr.kth_exact = nullptr;
            rr.idx_list = fail_list + static_cast<size_t>(sl) * R;
            rr.limit = retry_n + sl;
            rr.tau0 = rs.tau0;
            FR_CUDA(fr::launch_rescore(rr));
        }
    }

    // safety net: both launches return immediately when every query was certified
    fr::ScanArgs sa{};
    sa.corpus = ix->corpus;
    sa.keys_or_null = ms.keys_or_null;
    sa.queries = q;
    sa.n_rows = ix->rows;
    sa.dim = ix->dim;
    sa.bf16 = !f32_rows;  // the re-scan reads the rows the collection stores
    sa.l2 = l2;
    sa.k = k;
    sa.nq_total = B;
    sa.stream = s;
    if (!fr::scan_stream_fallback_serves(sa))
        return fail(FR_EUNSUP, "internal: no re-scan kernel for width %d", ix->dim);
    sa.grid = fr::scan_stream_fallback_grid(sa, ix->sm_count);
    FR_CUDA(ix->fb_partials.need(static_cast<size_t>(sa.grid) * B * k * sizeof(uint64_t)));
    sa.partials = static_cast<uint64_t *>(ix->fb_partials.p);
    FR_CUDA(fr::launch_scan_stream_fallback(sa, fail_count2, fail_list2));
    fr::MergeArgs mf{};
    mf.packed = sa.partials;
    mf.P = sa.grid;
    mf.shard_stride = static_cast<int64_t>(B) * k;
    mf.B = B;
    mf.k = k;
    mf.shards = false;
    mf.row_keys = ix->keys;
    mf.l2 = l2;
    mf.out_dist = d_out_dist;
    mf.out_packed = d_out_packed;
    mf.out_keys = d_out_keys;
    mf.only_flagged = static_cast<const uint8_t *>(ix->flags.p);
    mf.stream = s;
    FR_CUDA(fr::launch_merge_topk(mf));
    return FR_OK;
}

// Core of every search entry point: device queries in, merged lists out, all on stream `s`.
int search_on_stream(fr_index *ix, const float *d_queries, int B, int k, float *d_out_dist,
                     uint64_t *d_out_packed, int64_t *d_out_keys, cudaStream_t s) {
    if (B == 0) return FR_OK;
    // very large batches go through in slices so the per-call scratch (partial lists: streams x B x k' x 8 B)
    // stays bounded; a slice is still 16 corpus passes of 512 queries
    int MAX_SLICE = 8192;
    {
        const int ms_ = mma_slice(ix, k);  // widths other than 384 go through K2s in slices it can hold
        if (ms_ > 0 && ms_ < MAX_SLICE && ix->path != FR_PATH_STREAM) MAX_SLICE = ms_;
    }
    if (B > MAX_SLICE) {
        for (int b0 = 0; b0 < B; b0 += MAX_SLICE) {
            const int nb = B - b0 < MAX_SLICE ? B - b0 : MAX_SLICE;
            const size_t o = static_cast<size_t>(b0) * k;
            int rc = search_on_stream(ix, d_queries + static_cast<size_t>(b0) * ix->dim, nb, k,
                                      d_out_dist ? d_out_dist + o : nullptr, d_out_packed ? d_out_packed + o : nullptr,
                                      d_out_keys + o, s);
            if (rc != FR_OK) return rc;
        }
        return FR_OK;
    }
    ix->n_searches += 1;
    ix->n_queries += B;
    const bool eligible = mma_eligible(ix, k);
    if (ix->path == FR_PATH_MMA && !eligible)
        return fail(FR_EUNSUP,
                    "FR_PATH_MMA serves cosine and inner-product collections of width 384 (bf16 or fp32 rows, k <= 100) or 768 (bf16) and l2 collections of width 384 (bf16), with at least one row "
                    "(this one: dtype %d, dim %d, metric %d, k %d, rows %lld)",
                    ix->dtype, ix->dim, ix->metric, k, (long long)ix->rows);
    const bool k2s = eligible && small_serves(ix, B, fr::scan_mma_ksel(k, ix->mma_wide_lists));
    // small collections are launch-bound, not bandwidth-bound: K1 is 3 launches, the tensor-core path 11+
    // (scripts/latency_small.py: batch 1 over 1M rows 148 vs 165 us, over 10k rows 25 vs 53 us)
    const bool launch_bound = (B == 1 && ix->rows <= ix->small_rows_b1) || (B <= 4 && ix->rows <= ix->small_rows_b4);
    bool use_mma = eligible && (ix->path == FR_PATH_MMA ||
                                      (ix->path == FR_PATH_AUTO && !launch_bound && (B >= ix->mma_min_batch || k2s)));
    if (use_mma && ix->dtype == FR_F32) {
        // fp32 rows: the tensor-core scans need their bf16 copy.  If it cannot be allocated (a shard that fills the
        // HBM on its own) FR_PATH_AUTO stops asking for it and stays on the stream kernel.
        const int rs = ensure_shadow(ix, s);
        if (rs != FR_OK) {
            if (ix->path != FR_PATH_AUTO) return rs;
            ix->mma_f32_shadow = 0;
            use_mma = false;
        }
    }
    if (use_mma) return search_mma(ix, d_queries, B, k, d_out_dist, d_out_packed, d_out_keys, s);
    const size_t qbytes = static_cast<size_t>(B) * ix->dim * sizeof(float);
    const float *q = d_queries;
    if (ix->metric == FR_COSINE) {
        FR_CUDA(ix->q_prep.need(qbytes));
        FR_CUDA(ix->q_keys.need(static_cast<size_t>(B) * sizeof(int64_t)));
        fr::IngestArgs ia{};
        ia.src = d_queries;
        ia.n = B;
        ia.dim = ix->dim;
        ia.normalize = true;
        ia.bf16 = false;
        ia.corpus = static_cast<uint8_t *>(ix->q_prep.p);
        ia.keys = static_cast<int64_t *>(ix->q_keys.p);
        ia.stream = s;
        FR_CUDA(fr::launch_ingest(ia));
        q = static_cast<const float *>(ix->q_prep.p);
    }
    return search_stream(ix, q, B, k, d_out_dist, d_out_packed, d_out_keys, s);
}

// Everything a captured search graph bakes in: array addresses, row counts, routing options.
uint64_t state_hash(const fr_index *ix) {
    uint64_t h = 1469598103934665603ull;
    auto mix = [&h](uint64_t v) {
        h ^= v;
        h *= 1099511628211ull;
    };
    const DevBuf *bufs[] = {&ix->q_raw, &ix->q_prep, &ix->q_keys, &ix->partials, &ix->out_dist, &ix->out_keys, &ix->q_bf16,
                            &ix->err_bound, &ix->sel, &ix->sel_keys, &ix->flags, &ix->fail, &ix->fb_partials, &ix->tau,
                            &ix->progress, &ix->s_lists, &ix->hist, &ix->cmax, &ix->norm2, &ix->stats, &ix->kth_exact, &ix->r_q, &ix->r_misc, &ix->r_tau, &ix->r_partials, &ix->r_sel,
                            &ix->r_sel_keys};
    for (const DevBuf *b : bufs) mix(reinterpret_cast<uintptr_t>(b->p));
    mix(reinterpret_cast<uintptr_t>(ix->fail_mirror));
    mix(reinterpret_cast<uintptr_t>(ix->corpus));
    mix(reinterpret_cast<uintptr_t>(ix->shadow));
    mix(static_cast<uint64_t>(ix->shadow_rows) * 2u + (ix->shadow_dirty ? 1u : 0u));
    mix(reinterpret_cast<uintptr_t>(ix->keys));
    mix(static_cast<uint64_t>(ix->rows));
    mix(static_cast<uint64_t>(ix->cmax_rows));
    mix(ix->n_deleted > 0 ? 1u : 0u);
    for (int v : {ix->path, ix->mma_min_batch, ix->mma_small_max, ix->mma_co_groups, ix->mma_split, ix->mma_split_max,
                  ix->mma_debug, ix->mma_bound_scale_pct, ix->mma_max_lead, ix->mma_wide_lists, ix->mma_f32_shadow,
                  ix->retry_blocks_cur, ix->mma_score_hist})
        mix(static_cast<uint64_t>(static_cast<int64_t>(v)));
    mix(static_cast<uint64_t>(ix->small_rows_b1));
    mix(static_cast<uint64_t>(ix->small_rows_b4));
    return h;
}

void drop_graphs(fr_index *ix) {
    for (auto &kv : ix->graphs)
        if (kv.second.exec) cudaGraphExecDestroy(kv.second.exec);
    ix->graphs.clear();
}

int check_search_args(fr_index *ix, const void *q, int B, int k, const void *o1, const void *o2) {
    if (!ix) return fail(FR_EINVAL, "index is NULL");
    if (B < 0) return fail(FR_EINVAL, "B = %d is negative", B);
    if (k < 1) return fail(FR_EINVAL, "k = %d must be >= 1", k);
    if (k > FR_MAX_K) return fail(FR_EUNSUP, "k = %d exceeds FR_MAX_K = %d", k, FR_MAX_K);
    if (B > 0 && (!q || !o1 || !o2)) return fail(FR_EINVAL, "NULL buffer");
    return FR_OK;
}

// Host forms of the fusion / aggregation kernels: grow-only staging per device (a store object is constructed per
// request, rag_backend.py:611-643 -- allocating device memory and a stream per call cost more than the kernel).
struct FuseScratch {
    std::mutex mu;
    DevBuf in_a, in_b, out_a, out_b;
    cudaStream_t s = nullptr;
};
FuseScratch g_fuse[64];

// in_a / in_b: host inputs (in_b may be NULL) -> device; run(d_in_a, d_in_b, d_out_a, d_out_b, stream); outputs back.
template <typename Run>
int fuse_on_device(int device, const void *in_a, size_t in_a_bytes, const void *in_b, size_t in_b_bytes, void *out_a,
                          void *out_b, size_t out_bytes, Run run) {
    int rc = check_device(device, nullptr);
    if (rc != FR_OK) return rc;
    DeviceGuard g(device);
    FuseScratch &fs = g_fuse[device];
    std::lock_guard<std::mutex> lk(fs.mu);
    if (!fs.s) FR_CUDA(cudaStreamCreateWithFlags(&fs.s, cudaStreamNonBlocking));
    FR_CUDA(fs.in_a.need(in_a_bytes));
    FR_CUDA(fs.in_b.need(in_b_bytes ? in_b_bytes : 8));
    FR_CUDA(fs.out_a.need(out_bytes));
    FR_CUDA(fs.out_b.need(out_bytes));
    FR_CUDA(cudaMemcpyAsync(fs.in_a.p, in_a, in_a_bytes, cudaMemcpyHostToDevice, fs.s));
    if (in_b) FR_CUDA(cudaMemcpyAsync(fs.in_b.p, in_b, in_b_bytes, cudaMemcpyHostToDevice, fs.s));
    rc = run(fs.in_a.p, fs.in_b.p, fs.out_a.p, fs.out_b.p, fs.s);
    if (rc != FR_OK) return rc;
    FR_CUDA(cudaMemcpyAsync(out_a, fs.out_a.p, out_bytes, cudaMemcpyDeviceToHost, fs.s));
    FR_CUDA(cudaMemcpyAsync(out_b, fs.out_b.p, out_bytes, cudaMemcpyDeviceToHost, fs.s));
    FR_CUDA(cudaStreamSynchronize(fs.s));
    return FR_OK;
}

}  // namespace

extern "C" {

int fr_abi_version(void) { return FR_ABI_VERSION; }
const char *fr_last_error(void) { return g_last_error.c_str(); }
int64_t fr_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

int fr_index_create(int dim, int metric, int dtype, int device, int64_t reserve_rows, fr_index **out) {
    if (!out) return fail(FR_EINVAL, "out is NULL");
    *out = nullptr;
    if (dim < 8 || dim > FR_MAX_DIM || dim % 8 != 0)
        return fail(FR_EINVAL, "dim = %d must be a multiple of 8 in [8, %d]", dim, FR_MAX_DIM);
    if (metric != FR_COSINE && metric != FR_L2 && metric != FR_IP) return fail(FR_EINVAL, "unknown metric %d", metric);
    if (dtype != FR_BF16 && dtype != FR_F32) return fail(FR_EINVAL, "unknown dtype %d", dtype);
    if (reserve_rows < 0) return fail(FR_EINVAL, "reserve_rows is negative");
    int sm = 0;
    int rc = check_device(device, &sm);
    if (rc != FR_OK) return rc;
    DeviceGuard g(device);
    if (!g.ok) return fail(FR_ECUDA, "cudaSetDevice(%d) failed", device);
    fr_index *ix = new (std::nothrow) fr_index();
    if (!ix) return fail(FR_ENOMEM, "host allocation failed");
    ix->dim = dim;
    ix->metric = metric;
    ix->dtype = dtype;
    ix->device = device;
    ix->sm_count = sm;
    cudaError_t e = cudaStreamCreateWithFlags(&ix->stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ix->last_use, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventRecord(ix->last_use, ix->stream);
    if (e != cudaSuccess) {
        delete ix;
        return fail(FR_ECUDA, "stream/event creation failed: %s", cudaGetErrorString(e));
    }
    rc = grow(ix, reserve_rows > 0 ? reserve_rows : 1024, true);
    if (rc != FR_OK) {
        fr_index_destroy(ix);
        return rc;
    }
    *out = ix;
    return FR_OK;
}

int fr_index_destroy(fr_index *ix) {
    if (!ix) return FR_OK;
    {
        DeviceGuard g(ix->device);
        cudaDeviceSynchronize();
        drop_graphs(ix);
        if (ix->shadow) cudaFree(ix->shadow);
        if (ix->corpus) cudaFree(ix->corpus);
        if (ix->keys) cudaFree(ix->keys);
        DevBuf *bufs[] = {&ix->q_raw, &ix->q_prep, &ix->q_keys, &ix->partials, &ix->out_dist,
                          &ix->out_keys, &ix->stage_vecs, &ix->stage_keys, &ix->stage_rows,
                          &ix->q_bf16, &ix->err_bound, &ix->sel, &ix->sel_keys, &ix->flags, &ix->fail,
                          &ix->fb_partials, &ix->tau, &ix->progress, &ix->s_lists, &ix->cmax, &ix->norm2, &ix->stats, &ix->kth_exact, &ix->r_q, &ix->r_misc, &ix->r_tau,
                          &ix->r_partials, &ix->r_sel, &ix->r_sel_keys};
        for (DevBuf *b : bufs) b->release();
        for (auto &sl : ix->slots) {
            sl.pin.release();
            if (sl.done) cudaEventDestroy(sl.done);
        }
        if (ix->fail_mirror) cudaFreeHost(ix->fail_mirror);
        for (auto &pr : ix->prof_events) {
            cudaEventDestroy(pr.first);
            cudaEventDestroy(pr.second);
        }
        for (cudaEvent_t ev : ix->prof_pool) cudaEventDestroy(ev);
        if (ix->last_use) cudaEventDestroy(ix->last_use);
        if (ix->stream) cudaStreamDestroy(ix->stream);
    }
    delete ix;
    return FR_OK;
}

int fr_index_reserve(fr_index *ix, int64_t rows) {
    if (!ix) return fail(FR_EINVAL, "index is NULL");
    std::lock_guard<std::mutex> lk(ix->mu);
    DeviceGuard g(ix->device);
    return grow(ix, rows, true);
}

int fr_index_set_option(fr_index *ix, const char *name, int64_t value) {
    if (!ix || !name) return fail(FR_EINVAL, "NULL argument");
    std::lock_guard<std::mutex> lk(ix->mu);
    if (std::strcmp(name, "mma_min_batch") == 0) {
        if (value < 1) return fail(FR_EINVAL, "mma_min_batch must be >= 1");
        ix->mma_min_batch = static_cast<int>(value);
        return FR_OK;
    }
    if (std::strcmp(name, "mma_co_groups") == 0) {
        if (value < 1 || value > 8) return fail(FR_EINVAL, "mma_co_groups must be in [1, 8]");
        ix->mma_co_groups = static_cast<int>(value);
        return FR_OK;
    }
    if (std::strcmp(name, "mma_f32_shadow") == 0) {
        ix->mma_f32_shadow = value != 0;
        return FR_OK;
    }
    if (std::strcmp(name, "mma_wide_lists") == 0) {
        ix->mma_wide_lists = value != 0;
        return FR_OK;
    }
    if (std::strcmp(name, "mma_max_lead") == 0) {
        if (value < 0 || value > 1024) return fail(FR_EINVAL, "mma_max_lead must be in [0, 1024]");
        ix->mma_max_lead = static_cast<int>(value);
        return FR_OK;
    }
    if (std::strcmp(name, "mma_bound_scale_pct") == 0) {
        if (value < 100 || value > 100000) return fail(FR_EINVAL, "mma_bound_scale_pct must be in [100, 100000]");
        ix->mma_bound_scale_pct = static_cast<int>(value);
        return FR_OK;
    }
    if (std::strcmp(name, "use_graphs") == 0) {
        ix->use_graphs = value != 0;
        return FR_OK;
    }
    if (std::strcmp(name, "graph_max_bytes") == 0) {
        ix->graph_max_bytes = value;
        return FR_OK;
    }
    if (std::strcmp(name, "host_debug") == 0) {
        ix->host_debug = static_cast<int>(value);
        return FR_OK;
    }
    if (std::strcmp(name, "small_rows_b1") == 0) {
        ix->small_rows_b1 = value;
        return FR_OK;
    }
    if (std::strcmp(name, "small_rows_b4") == 0) {
        ix->small_rows_b4 = value;
        return FR_OK;
    }
    if (std::strcmp(name, "mma_split") == 0) {
        if (value < -1 || value > 1) return fail(FR_EINVAL, "mma_split must be -1 (auto), 0 or 1");
        ix->mma_split = static_cast<int>(value);
        return FR_OK;
    }
    if (std::strcmp(name, "mma_split_max") == 0) {
        if (value < 0 || value > 64) return fail(FR_EINVAL, "mma_split_max must be in [0, 64]");
        ix->mma_split_max = static_cast<int>(value);
        return FR_OK;
    }
    if (std::strcmp(name, "mma_small_max") == 0) {
        if (value < 0 || value > 64) return fail(FR_EINVAL, "mma_small_max must be in [0, 64]");
        ix->mma_small_max = static_cast<int>(value);
        return FR_OK;
    }
    if (std::strcmp(name, "mma_retry_blocks") == 0) {
        if (value < -1 || value > 4096) return fail(FR_EINVAL, "mma_retry_blocks must be -1 (one per 128 queries), 0 (adaptive) or a count");
        ix->mma_retry_blocks = static_cast<int>(value);
        return FR_OK;
    }
    if (std::strcmp(name, "mma_debug") == 0) {
        ix->mma_debug = static_cast<int>(value);
        return FR_OK;
    }
    if (std::strcmp(name, "mma_score_hist") == 0) {
        ix->mma_score_hist = value != 0;
        return FR_OK;
    }
    if (std::strcmp(name, "profile") == 0) {
        ix->profile = value != 0;
        return FR_OK;
    }
    if (std::strcmp(name, "path") == 0) {
        if (value < FR_PATH_AUTO || value > FR_PATH_MMA) return fail(FR_EINVAL, "path = %lld is not a FR_PATH_* value", (long long)value);
        ix->path = static_cast<int>(value);
        return FR_OK;
    }
    return fail(FR_EINVAL, "unknown option '%s'", name);
}

int fr_index_profile_read(fr_index *ix, double *out_scan_ms, int64_t *out_scan_launches, int64_t *out_searches) {
    if (!ix || !out_scan_ms || !out_scan_launches || !out_searches) return fail(FR_EINVAL, "NULL argument");
    std::lock_guard<std::mutex> lk(ix->mu);
    DeviceGuard g(ix->device);
    double total = 0.0;
    for (auto &pr : ix->prof_events) {
        FR_CUDA(cudaEventSynchronize(pr.second));
        float ms = 0.0f;
        FR_CUDA(cudaEventElapsedTime(&ms, pr.first, pr.second));
        total += ms;
        ix->prof_pool.push_back(pr.first);
        ix->prof_pool.push_back(pr.second);
    }
    *out_scan_ms = total;
    *out_scan_launches = ix->prof_launches;
    *out_searches = static_cast<int64_t>(ix->prof_events.size());
    ix->prof_events.clear();
    ix->prof_launches = 0;
    return FR_OK;
}

int fr_index_get_stat(fr_index *ix, const char *name, int64_t *out) {
    if (!ix || !name || !out) return fail(FR_EINVAL, "NULL argument");
    std::lock_guard<std::mutex> lk(ix->mu);
    if (std::strcmp(name, "searches") == 0) {
        *out = ix->n_searches;
        return FR_OK;
    }
    if (std::strcmp(name, "queries") == 0) {
        *out = ix->n_queries;
        return FR_OK;
    }
    if (std::strcmp(name, "mma_queries") == 0) {
        *out = ix->n_mma_queries;
        return FR_OK;
    }
    if (std::strcmp(name, "mma_retry_blocks") == 0) {
        *out = ix->retry_blocks_cur;
        return FR_OK;
    }
    if (std::strcmp(name, "graph_replays") == 0) {
        *out = ix->n_graph_replays;
        return FR_OK;
    }
    const bool unc = std::strcmp(name, "mma_uncertified_queries") == 0;
    if (unc || std::strcmp(name, "mma_rescanned_queries") == 0) {
        *out = 0;
        if (!ix->stats.p) return FR_OK;
        DeviceGuard g(ix->device);
        FR_CUDA(cudaDeviceSynchronize());
        unsigned long long v[2] = {0, 0};
        FR_CUDA(cudaMemcpy(v, ix->stats.p, sizeof(v), cudaMemcpyDeviceToHost));
        *out = static_cast<int64_t>(unc ? v[0] : v[1]);
        return FR_OK;
    }
    return fail(FR_EINVAL, "unknown stat '%s'", name);
}

int fr_index_count(fr_index *ix, int64_t *out) {
    if (!ix || !out) return fail(FR_EINVAL, "NULL argument");
    std::lock_guard<std::mutex> lk(ix->mu);
    *out = ix->rows - ix->n_deleted;
    return FR_OK;
}

int fr_index_rows(fr_index *ix, int64_t *out) {
    if (!ix || !out) return fail(FR_EINVAL, "NULL argument");
    std::lock_guard<std::mutex> lk(ix->mu);
    *out = ix->rows;
    return FR_OK;
}

int fr_index_upsert(fr_index *ix, const float *vecs, const int64_t *keys, int64_t n) {
    if (!ix) return fail(FR_EINVAL, "index is NULL");
    if (n < 0) return fail(FR_EINVAL, "n is negative");
    if (n == 0) return FR_OK;
    if (!vecs || !keys) return fail(FR_EINVAL, "NULL buffer");
    std::lock_guard<std::mutex> lk(ix->mu);
    DeviceGuard g(ix->device);
    int rc = rebuild_keymap(ix);
    if (rc != FR_OK) return rc;
    // resolve target rows on the host: existing key -> its row (in place), new key -> append
    std::vector<int64_t> target(static_cast<size_t>(n));
    std::unordered_map<int64_t, int64_t> last_writer;  // row -> index of the last vector aimed at it
    int64_t new_rows = ix->rows;
    for (int64_t i = 0; i < n; ++i)
        if (keys[i] == fr::KEY_TOMBSTONE || keys[i] == FR_KEY_NONE)
            return fail(FR_EINVAL, "key %lld is reserved (INT64_MIN marks deleted rows, -1 pads result lists)", (long long)keys[i]);
    for (int64_t i = 0; i < n; ++i) {
        auto it = ix->keymap.find(keys[i]);
        int64_t row;
        if (it != ix->keymap.end()) {
            row = it->second;
        } else {
            row = new_rows++;
            ix->keymap.emplace(keys[i], row);
        }
        target[i] = row;
        auto lw = last_writer.find(row);
        if (lw != last_writer.end()) {
            target[lw->second] = -1;  // duplicate key inside this call: the last one wins
            lw->second = i;
        } else {
            last_writer.emplace(row, i);
        }
    }
    if (new_rows > 0xfffffff0ll) {
        ix->keymap_valid = false;
        return fail(FR_EUNSUP, "a shard holds at most 2^32-16 rows");
    }
    if (static_cast<int64_t>(last_writer.size()) != new_rows - ix->rows) {  // an existing row is overwritten
        ix->shadow_dirty = true;
        ix->cmax_rows = 0;  // (the maximum is recomputed over all rows: an overwrite may have raised it)
    }
    rc = grow(ix, new_rows, false);
    if (rc != FR_OK) {
        ix->keymap_valid = false;  // the map now names rows that were never written
        return rc;
    }
    cudaStream_t s = ix->stream;
    rc = begin_use(ix, s);
    if (rc != FR_OK) {
        ix->keymap_valid = false;
        return rc;
    }
    const int64_t CH = 1 << 16;
    const size_t vb = static_cast<size_t>(ix->dim) * sizeof(float);
    auto write_chunks = [&]() -> int {
    for (int64_t lo = 0; lo < n; lo += CH) {
        const int64_t m = (n - lo < CH) ? (n - lo) : CH;
        FR_CUDA(ix->stage_vecs.need(static_cast<size_t>(m) * vb));
        FR_CUDA(ix->stage_keys.need(static_cast<size_t>(m) * sizeof(int64_t)));
        FR_CUDA(ix->stage_rows.need(static_cast<size_t>(m) * sizeof(int64_t)));
        FR_CUDA(cudaMemcpyAsync(ix->stage_vecs.p, vecs + lo * ix->dim, static_cast<size_t>(m) * vb, cudaMemcpyHostToDevice, s));
        FR_CUDA(cudaMemcpyAsync(ix->stage_keys.p, keys + lo, static_cast<size_t>(m) * sizeof(int64_t), cudaMemcpyHostToDevice, s));
        FR_CUDA(cudaMemcpyAsync(ix->stage_rows.p, target.data() + lo, static_cast<size_t>(m) * sizeof(int64_t), cudaMemcpyHostToDevice, s));
        fr::IngestArgs ia{};
        ia.src = static_cast<const float *>(ix->stage_vecs.p);
        ia.src_keys = static_cast<const int64_t *>(ix->stage_keys.p);
        ia.target_rows = static_cast<const int64_t *>(ix->stage_rows.p);
        ia.n = m;
        ia.dim = ix->dim;
        ia.normalize = ix->metric == FR_COSINE;
        ia.bf16 = ix->dtype == FR_BF16;
        ia.corpus = ix->corpus;
        ia.keys = ix->keys;
        ia.stream = s;
        FR_CUDA(fr::launch_ingest(ia));
        FR_CUDA(cudaStreamSynchronize(s));  // staging buffers are reused by the next chunk
    }
    return FR_OK;
    };
    rc = write_chunks();
    if (rc != FR_OK) {
        // the map already names rows that were never written: rebuild it from the device before the next use
        ix->keymap_valid = false;
        end_use(ix, s);
        return rc;
    }
    ix->rows = new_rows;
    return end_use(ix, s);
}

int fr_index_delete(fr_index *ix, const int64_t *keys, int64_t n, int64_t *out_deleted) {
    if (!ix) return fail(FR_EINVAL, "index is NULL");
    if (out_deleted) *out_deleted = 0;
    if (n < 0) return fail(FR_EINVAL, "n is negative");
    if (n == 0) return FR_OK;
    if (!keys) return fail(FR_EINVAL, "NULL buffer");
    std::lock_guard<std::mutex> lk(ix->mu);
    DeviceGuard g(ix->device);
    int rc = rebuild_keymap(ix);
    if (rc != FR_OK) return rc;
    std::vector<int64_t> rows;
    for (int64_t i = 0; i < n; ++i) {
        auto it = ix->keymap.find(keys[i]);
        if (it == ix->keymap.end()) continue;
        rows.push_back(it->second);
        ix->keymap.erase(it);
    }
    if (!rows.empty()) {
        cudaStream_t s = ix->stream;
        rc = begin_use(ix, s);
        if (rc != FR_OK) return rc;
        FR_CUDA(ix->stage_rows.need(rows.size() * sizeof(int64_t)));
        FR_CUDA(cudaMemcpyAsync(ix->stage_rows.p, rows.data(), rows.size() * sizeof(int64_t), cudaMemcpyHostToDevice, s));
        FR_CUDA(fr::launch_fill_keys(ix->keys, static_cast<const int64_t *>(ix->stage_rows.p),
                                     static_cast<int64_t>(rows.size()), fr::KEY_TOMBSTONE, s));
        FR_CUDA(cudaStreamSynchronize(s));
        ix->n_deleted += static_cast<int64_t>(rows.size());
        rc = end_use(ix, s);
        if (rc != FR_OK) return rc;
    }
    if (out_deleted) *out_deleted = static_cast<int64_t>(rows.size());
    return FR_OK;
}

int fr_index_append_device(fr_index *ix, const float *d_vecs, const int64_t *d_keys, int64_t first_key, int64_t n,
                           void *stream) {
    if (!ix) return fail(FR_EINVAL, "index is NULL");
    if (n < 0) return fail(FR_EINVAL, "n is negative");
    if (n == 0) return FR_OK;
    if (!d_vecs) return fail(FR_EINVAL, "NULL buffer");
    std::lock_guard<std::mutex> lk(ix->mu);
    DeviceGuard g(ix->device);
    if (ix->rows + n > 0xfffffff0ll) return fail(FR_EUNSUP, "a shard holds at most 2^32-16 rows");
    int rc = grow(ix, ix->rows + n, false);
    if (rc != FR_OK) return rc;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    rc = begin_use(ix, s);
    if (rc != FR_OK) return rc;
    fr::IngestArgs ia{};
    ia.src = d_vecs;
    ia.src_keys = d_keys;
    ia.first_key = first_key;
    ia.base_row = ix->rows;
    ia.n = n;
    ia.dim = ix->dim;
    ia.normalize = ix->metric == FR_COSINE;
    ia.bf16 = ix->dtype == FR_BF16;
    ia.corpus = ix->corpus;
    ia.keys = ix->keys;
    ia.stream = s;
    FR_CUDA(fr::launch_ingest(ia));
    ix->rows += n;
    ix->keymap_valid = false;  // rebuilt lazily from the device keys if upsert/delete is used later
    ix->keymap.clear();
    return end_use(ix, s);
}

int fr_index_get_rows(fr_index *ix, int64_t first_row, int64_t n, float *out_vecs, int64_t *out_keys) {
    if (!ix) return fail(FR_EINVAL, "index is NULL");
    std::lock_guard<std::mutex> lk(ix->mu);
    if (first_row < 0 || n < 0 || first_row + n > ix->rows)
        return fail(FR_EINVAL, "rows [%lld, %lld) out of range (index holds %lld)", (long long)first_row,
                    (long long)(first_row + n), (long long)ix->rows);
    if (n == 0) return FR_OK;
    DeviceGuard g(ix->device);
    FR_CUDA(cudaDeviceSynchronize());
    const size_t rb = ix->row_bytes();
    if (out_vecs) {
        if (ix->dtype == FR_F32) {
            FR_CUDA(cudaMemcpy(out_vecs, ix->corpus + first_row * rb, static_cast<size_t>(n) * rb, cudaMemcpyDeviceToHost));
        } else {
            std::vector<uint16_t> tmp(static_cast<size_t>(n) * ix->dim);
            FR_CUDA(cudaMemcpy(tmp.data(), ix->corpus + first_row * rb, tmp.size() * 2, cudaMemcpyDeviceToHost));
            for (size_t i = 0; i < tmp.size(); ++i) {
                const uint32_t u = static_cast<uint32_t>(tmp[i]) << 16;
                std::memcpy(&out_vecs[i], &u, 4);
            }
        }
    }
    if (out_keys)
        FR_CUDA(cudaMemcpy(out_keys, ix->keys + first_row, static_cast<size_t>(n) * sizeof(int64_t), cudaMemcpyDeviceToHost));
    return FR_OK;
}

int fr_index_export_raw(fr_index *ix, int64_t first_row, int64_t n, void *out_rows, int64_t *out_keys) {
    if (!ix) return fail(FR_EINVAL, "index is NULL");
    std::lock_guard<std::mutex> lk(ix->mu);
    if (first_row < 0 || n < 0 || first_row + n > ix->rows)
        return fail(FR_EINVAL, "rows [%lld, %lld) out of range (index holds %lld)", (long long)first_row,
                    (long long)(first_row + n), (long long)ix->rows);
    if (n == 0) return FR_OK;
    DeviceGuard g(ix->device);
    FR_CUDA(cudaDeviceSynchronize());
    const size_t rb = ix->row_bytes();
    if (out_rows)
        FR_CUDA(cudaMemcpy(out_rows, ix->corpus + first_row * rb, static_cast<size_t>(n) * rb, cudaMemcpyDeviceToHost));
    if (out_keys)
        FR_CUDA(cudaMemcpy(out_keys, ix->keys + first_row, static_cast<size_t>(n) * sizeof(int64_t), cudaMemcpyDeviceToHost));
    return FR_OK;
}

int fr_index_import_raw(fr_index *ix, const void *rows, const int64_t *keys, int64_t n) {
    if (!ix) return fail(FR_EINVAL, "index is NULL");
    if (n < 0) return fail(FR_EINVAL, "n is negative");
    if (n == 0) return FR_OK;
    if (!rows || !keys) return fail(FR_EINVAL, "NULL buffer");
    std::lock_guard<std::mutex> lk(ix->mu);
    DeviceGuard g(ix->device);
    if (ix->rows + n > 0xfffffff0ll) return fail(FR_EUNSUP, "a shard holds at most 2^32-16 rows");
    for (int64_t i = 0; i < n; ++i)
        if (keys[i] == FR_KEY_NONE) return fail(FR_EINVAL, "key -1 is reserved (it pads result lists)");
    int rc = grow(ix, ix->rows + n, true);
    if (rc != FR_OK) return rc;
    cudaStream_t s = ix->stream;
    rc = begin_use(ix, s);
    if (rc != FR_OK) return rc;
    const size_t rb = ix->row_bytes();
    // the source is typically a memory-mapped shard file: copy in bounded pieces so the driver's
    // staging of pageable memory stays small
    const int64_t CH = static_cast<int64_t>((size_t(64) << 20) / rb);
    for (int64_t lo = 0; lo < n; lo += CH) {
        const int64_t m = (n - lo < CH) ? (n - lo) : CH;
        FR_CUDA(cudaMemcpyAsync(ix->corpus + static_cast<size_t>(ix->rows + lo) * rb,
                                static_cast<const uint8_t *>(rows) + static_cast<size_t>(lo) * rb,
                                static_cast<size_t>(m) * rb, cudaMemcpyHostToDevice, s));
        FR_CUDA(cudaStreamSynchronize(s));
    }
    FR_CUDA(cudaMemcpyAsync(ix->keys + ix->rows, keys, static_cast<size_t>(n) * sizeof(int64_t), cudaMemcpyHostToDevice, s));
    FR_CUDA(cudaStreamSynchronize(s));
    for (int64_t i = 0; i < n; ++i)
        if (keys[i] == fr::KEY_TOMBSTONE) ix->n_deleted += 1;
    ix->rows += n;
    ix->keymap_valid = false;
    ix->keymap.clear();
    return end_use(ix, s);
}

int fr_index_lookup_rows(fr_index *ix, const int64_t *keys, int64_t n, int64_t *out_rows) {
    if (!ix) return fail(FR_EINVAL, "index is NULL");
    if (n < 0) return fail(FR_EINVAL, "n is negative");
    if (n == 0) return FR_OK;
    if (!keys || !out_rows) return fail(FR_EINVAL, "NULL buffer");
    std::lock_guard<std::mutex> lk(ix->mu);
    DeviceGuard g(ix->device);
    int rc = rebuild_keymap(ix);
    if (rc != FR_OK) return rc;
    for (int64_t i = 0; i < n; ++i) {
        auto it = ix->keymap.find(keys[i]);
        out_rows[i] = it == ix->keymap.end() ? -1 : it->second;
    }
    return FR_OK;
}

namespace {
// a pinned staging slot of one host search, held from before the enqueue until the results are copied out
struct SlotLease {
    fr_index *ix;
    int id = -1;
    explicit SlotLease(fr_index *ix_) : ix(ix_) {
        std::unique_lock<std::mutex> lk(ix->slot_mu);
        ix->slot_cv.wait(lk, [&] {
            for (int i = 0; i < fr_index::N_SLOTS; ++i)
                if (!ix->slots[i].busy) return true;
            return false;
        });
        for (int i = 0; i < fr_index::N_SLOTS; ++i)
            if (!ix->slots[i].busy) {
                ix->slots[i].busy = true;
                id = i;
                break;
            }
    }
    ~SlotLease() {
        {
            std::lock_guard<std::mutex> lk(ix->slot_mu);
            ix->slots[id].busy = false;
        }
        ix->slot_cv.notify_one();
    }
    fr_index::HostSlot &slot() { return ix->slots[id]; }
};
}  // namespace

int fr_index_search(fr_index *ix, const float *queries, int B, int k, float *out_dist, int64_t *out_keys) {
    int rc = check_search_args(ix, queries, B, k, out_dist, out_keys);
    if (rc != FR_OK || B == 0) return rc;
    SlotLease lease(ix);
    fr_index::HostSlot &slot = lease.slot();
    DeviceGuard g(ix->device);
    cudaStream_t s = ix->stream;
    const size_t qb = static_cast<size_t>(B) * ix->dim * sizeof(float);
    const size_t db = static_cast<size_t>(B) * k * sizeof(float);
    const size_t kb = static_cast<size_t>(B) * k * sizeof(int64_t);
    const size_t db_al = (db + 15) & ~static_cast<size_t>(15), kb_al = (kb + 15) & ~static_cast<size_t>(15);
    // The slot is ours alone and its previous search has been waited for, but growing it frees and allocates pinned
    // memory, which synchronises the device -- while another thread may be capturing this shard's stream into a graph
    // (relaxed capture mode checks nothing).  So the rare growth and the event's creation take the shard lock, under
    // which every capture runs; the common case (the block is large enough) takes none.
    if (!slot.done || slot.pin.bytes < qb + db_al + kb_al) {
        std::lock_guard<std::mutex> lk(ix->mu);
        if (!slot.done) FR_CUDA(cudaEventCreateWithFlags(&slot.done, cudaEventDisableTiming));
        FR_CUDA(slot.pin.need(qb + db_al + kb_al));
    }
    uint8_t *pin = static_cast<uint8_t *>(slot.pin.p);
    // every block 16-byte aligned: the kernels may read the queries as float4 straight from the pinned block
    uint8_t *pin_keys = pin, *pin_dist = pin + kb_al, *pin_q = pin + kb_al + db_al;
    std::memcpy(pin_q, queries, qb);
    {
    std::unique_lock<std::mutex> lk(ix->mu);
    FR_CUDA(ix->q_raw.need(qb));
    FR_CUDA(ix->out_dist.need(db));
    FR_CUDA(ix->out_keys.need(kb));
    // the whole call as it is enqueued on `s`: H2D of the queries, the search, D2H of the results
    // Small batches of a cosine collection skip the three copy operations: the query-preparation kernel reads the
    // pinned block directly (the only reader of the raw queries) and the last kernel of the chain writes its
    // B x k results straight into it -- each copy node costs more than the kernels around it at this size.
    const bool zero_copy = ix->metric == FR_COSINE && B <= 64;
    auto enqueue = [&]() -> int {
        if (zero_copy)
            return search_on_stream(ix, reinterpret_cast<const float *>(pin_q), B, k, reinterpret_cast<float *>(pin_dist),
                                    nullptr, reinterpret_cast<int64_t *>(pin_keys), s);
        FR_CUDA(cudaMemcpyAsync(ix->q_raw.p, pin_q, qb, cudaMemcpyHostToDevice, s));
        int r = search_on_stream(ix, static_cast<const float *>(ix->q_raw.p), B, k, static_cast<float *>(ix->out_dist.p),
                                 nullptr, static_cast<int64_t *>(ix->out_keys.p), s);
        if (r != FR_OK) return r;
        FR_CUDA(cudaMemcpyAsync(pin_dist, ix->out_dist.p, db, cudaMemcpyDeviceToHost, s));
        FR_CUDA(cudaMemcpyAsync(pin_keys, ix->out_keys.p, kb, cudaMemcpyDeviceToHost, s));
        return FR_OK;
    };
    // Small collections: replay a captured graph of the call (same shape, same staging slot, nothing it baked in
    // changed).  The first call of a shape runs eagerly (it sizes the scratch), the second is captured, later ones replay.
    bool graphable = ix->use_graphs && !ix->profile && B <= 8192 &&
                     static_cast<int64_t>(ix->rows) * static_cast<int64_t>(ix->row_bytes()) <= ix->graph_max_bytes;
    // graph work only while no other caller waits for its results (fr_host.h: graph_wait_mutex); eager otherwise
    std::unique_lock<std::shared_mutex> graph_lk(fr::graph_wait_mutex(), std::defer_lock);
    if (graphable && !(ix->host_debug & 4)) graphable = graph_lk.try_lock();
    fr_index::SearchGraph *sg = nullptr;
    auto hash_now = [&]() {  // what a graph bakes in: the shard's state and this slot's pinned block
        uint64_t h = state_hash(ix);
        h ^= reinterpret_cast<uintptr_t>(pin);
        h *= 1099511628211ull;
        return h;
    };
    if (graphable) {
        if (ix->graphs.size() > 64) drop_graphs(ix);
        plan_retry_blocks(ix);  // part of what a graph bakes in
        sg = &ix->graphs[(static_cast<uint64_t>(static_cast<uint32_t>(B)) << 32) | (static_cast<uint64_t>(static_cast<uint32_t>(k)) << 4) |
                         static_cast<uint64_t>(lease.id)];
        const uint64_t h = hash_now();
        if (sg->exec && sg->state != h) {
            cudaGraphExecDestroy(sg->exec);
            sg->exec = nullptr;
        }
        if (!sg->exec && !sg->failed && sg->seen == h) {  // second stable call of this shape: capture it
            const int64_t l0 = fr_launch_count(), s0 = ix->n_searches, q0 = ix->n_queries, m0 = ix->n_mma_queries;
            cudaGraph_t graph = nullptr;
            bool ok = cudaStreamBeginCapture(s, (ix->host_debug & 1) ? cudaStreamCaptureModeThreadLocal : cudaStreamCaptureModeRelaxed) == cudaSuccess;
            if (ok) {
                const int r = enqueue();
                ok = cudaStreamEndCapture(s, &graph) == cudaSuccess && r == FR_OK && graph != nullptr;
            }
            // capture enqueues nothing: take back what the bookkeeping counted
            sg->launches = fr_launch_count() - l0;
            sg->d_searches = ix->n_searches - s0;
            sg->d_queries = ix->n_queries - q0;
            sg->d_mma_queries = ix->n_mma_queries - m0;
            g_launches.fetch_sub(sg->launches, std::memory_order_relaxed);
            ix->n_searches = s0;
            ix->n_queries = q0;
            ix->n_mma_queries = m0;
            if (ok) ok = hash_now() == h;  // the capture must not have moved any scratch
            if (ok) ok = cudaGraphInstantiate(&sg->exec, graph, 0) == cudaSuccess;
            if (graph) cudaGraphDestroy(graph);
            if (!ok) {
                cudaGetLastError();
                sg->exec = nullptr;
                sg->failed = true;
            } else {
                sg->state = h;
            }
        }
    }
    rc = begin_use(ix, s);
    if (rc != FR_OK) return rc;
    if (sg && sg->exec) {
        FR_CUDA(cudaGraphLaunch(sg->exec, s));
        g_launches.fetch_add(sg->launches, std::memory_order_relaxed);
        ix->n_searches += sg->d_searches;
        ix->n_queries += sg->d_queries;
        ix->n_mma_queries += sg->d_mma_queries;
        ix->n_graph_replays += 1;
    } else {
        rc = enqueue();
        if (rc != FR_OK) {
            cudaStreamSynchronize(s);  // whatever part of the chain was enqueued may still write the slot
            return rc;
        }
        if (sg) sg->seen = hash_now();
    }
    if (graph_lk.owns_lock()) graph_lk.unlock();
    rc = end_use(ix, s);
    if (rc != FR_OK) return rc;
    FR_CUDA(cudaEventRecord(slot.done, s));
    if (ix->host_debug & 2) FR_CUDA(cudaEventSynchronize(slot.done));
    }  // the shard is free for the next caller's enqueue; this call waits for its own results only
    {
        std::shared_lock<std::shared_mutex> wait_lk(fr::graph_wait_mutex(), std::defer_lock);
        if (!(ix->host_debug & 4)) wait_lk.lock();
        FR_CUDA(cudaEventSynchronize(slot.done));
    }
    std::memcpy(out_dist, pin_dist, db);
    std::memcpy(out_keys, pin_keys, kb);
    return FR_OK;
}
int fr_index_search_device(fr_index *ix, const float *d_queries, int B, int k, float *d_out_dist, int64_t *d_out_keys,
                           void *stream) {
    int rc = check_search_args(ix, d_queries, B, k, d_out_dist, d_out_keys);
    if (rc != FR_OK || B == 0) return rc;
    std::lock_guard<std::mutex> lk(ix->mu);
    DeviceGuard g(ix->device);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    rc = begin_use(ix, s);
    if (rc != FR_OK) return rc;
    rc = search_on_stream(ix, d_queries, B, k, d_out_dist, nullptr, d_out_keys, s);
    // recorded on failure too: part of the chain may be enqueued, and the next caller must wait for it
    const int rc2 = end_use(ix, s);
    return rc != FR_OK ? rc : rc2;
}

int fr_index_search_partial_device(fr_index *ix, const float *d_queries, int B, int k, uint64_t *d_out_packed,
                                   int64_t *d_out_keys, void *stream) {
    int rc = check_search_args(ix, d_queries, B, k, d_out_packed, d_out_keys);
    if (rc != FR_OK || B == 0) return rc;
    std::lock_guard<std::mutex> lk(ix->mu);
    DeviceGuard g(ix->device);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    rc = begin_use(ix, s);
    if (rc != FR_OK) return rc;
    rc = search_on_stream(ix, d_queries, B, k, nullptr, d_out_packed, d_out_keys, s);
    const int rc2 = end_use(ix, s);
    return rc != FR_OK ? rc : rc2;
}

int fr_merge_shards_device(int device, int metric, const uint64_t *d_packed, const int64_t *d_keys,
                           int64_t shard_stride_elems, int G, int B, int k, float *d_out_dist, int64_t *d_out_keys,
                           void *stream) {
    if (G < 1 || B < 0 || k < 1) return fail(FR_EINVAL, "bad G/B/k (%d/%d/%d)", G, B, k);
    if (k > FR_MAX_K) return fail(FR_EUNSUP, "k = %d exceeds FR_MAX_K = %d", k, FR_MAX_K);
    if (B == 0) return FR_OK;
    if (!d_packed || !d_keys || !d_out_dist || !d_out_keys) return fail(FR_EINVAL, "NULL buffer");
    if (shard_stride_elems < static_cast<int64_t>(B) * k) return fail(FR_EINVAL, "shard stride smaller than B*k");
    int rc = check_device(device, nullptr);
    if (rc != FR_OK) return rc;
    DeviceGuard g(device);
    fr::MergeArgs ma{};
    ma.packed = d_packed;
    ma.P = G;
    ma.shard_stride = shard_stride_elems;
    ma.B = B;
    ma.k = k;
    ma.shards = true;
    ma.shard_keys = d_keys;
    ma.l2 = metric == FR_L2;
    ma.out_dist = d_out_dist;
    ma.out_keys = d_out_keys;
    ma.stream = static_cast<cudaStream_t>(stream);
    FR_CUDA(fr::launch_merge_topk(ma));
    return FR_OK;
}

int fr_rrf_fuse_device(int device, const int64_t *d_keys, int L, int B, int kp, int k_rrf, int k_out,
                       double *d_out_score, int64_t *d_out_keys, void *stream) {
    if (L < 1 || B < 0 || kp < 1 || k_out < 1 || k_rrf < 0) return fail(FR_EINVAL, "bad L/B/kp/k_rrf/k_out");
    if (static_cast<int64_t>(L) * kp > 2048) return fail(FR_EUNSUP, "L*kp = %d exceeds 2048 candidates per query", L * kp);
    if (B == 0) return FR_OK;
    if (!d_keys || !d_out_score || !d_out_keys) return fail(FR_EINVAL, "NULL buffer");
    int rc = check_device(device, nullptr);
    if (rc != FR_OK) return rc;
    DeviceGuard g(device);
    fr::RrfArgs ra{d_keys, L, B, kp, k_rrf, k_out, d_out_score, d_out_keys, static_cast<cudaStream_t>(stream)};
    FR_CUDA(fr::launch_rrf_fuse(ra));
    return FR_OK;
}

int fr_rrf_fuse(int device, const int64_t *keys, int L, int B, int kp, int k_rrf, int k_out, double *out_score,
                int64_t *out_keys) {
    if (L < 1 || B < 0 || kp < 1 || k_out < 1 || k_rrf < 0) return fail(FR_EINVAL, "bad L/B/kp/k_rrf/k_out");
    if (B == 0) return FR_OK;
    if (!keys || !out_score || !out_keys) return fail(FR_EINVAL, "NULL buffer");
    return fuse_on_device(device, keys, static_cast<size_t>(L) * B * kp * sizeof(int64_t), nullptr, 0, out_score, out_keys,
                          static_cast<size_t>(B) * k_out * 8,
                          [&](void *d_in, void *, void *d_sc, void *d_k, cudaStream_t s) {
                              return fr_rrf_fuse_device(device, static_cast<const int64_t *>(d_in), L, B, kp, k_rrf, k_out,
                                                        static_cast<double *>(d_sc), static_cast<int64_t *>(d_k), s);
                          });
}

int fr_score_fuse_device(int device, const float *d_dist, const int64_t *d_keys, int L, int B, int kp, int k_out,
                         double *d_out_score, int64_t *d_out_keys, void *stream) {
    if (L < 1 || B < 0 || kp < 1 || k_out < 1) return fail(FR_EINVAL, "bad L/B/kp/k_out");
    if (static_cast<int64_t>(L) * kp > 2048 || L > 64)
        return fail(FR_EUNSUP, "L*kp = %d exceeds 2048 candidates per query (or L > 64)", L * kp);
    if (B == 0) return FR_OK;
    if (!d_dist || !d_keys || !d_out_score || !d_out_keys) return fail(FR_EINVAL, "NULL buffer");
    int rc = check_device(device, nullptr);
    if (rc != FR_OK) return rc;
    DeviceGuard g(device);
    fr::ScoreFuseArgs sa{d_dist, d_keys, L, B, kp, k_out, d_out_score, d_out_keys, static_cast<cudaStream_t>(stream)};
    FR_CUDA(fr::launch_score_fuse(sa));
    return FR_OK;
}

int fr_score_fuse(int device, const float *dist, const int64_t *keys, int L, int B, int kp, int k_out, double *out_score,
                  int64_t *out_keys) {
    if (L < 1 || B < 0 || kp < 1 || k_out < 1) return fail(FR_EINVAL, "bad L/B/kp/k_out");
    if (B == 0) return FR_OK;
    if (!dist || !keys || !out_score || !out_keys) return fail(FR_EINVAL, "NULL buffer");
    const size_t ne = static_cast<size_t>(L) * B * kp;
    return fuse_on_device(device, dist, ne * sizeof(float), keys, ne * sizeof(int64_t), out_score, out_keys,
                          static_cast<size_t>(B) * k_out * 8,
                          [&](void *d_d, void *d_k, void *d_sc, void *d_ok, cudaStream_t s) {
                              return fr_score_fuse_device(device, static_cast<const float *>(d_d),
                                                          static_cast<const int64_t *>(d_k), L, B, kp, k_out,
                                                          static_cast<double *>(d_sc), static_cast<int64_t *>(d_ok), s);
                          });
}

int fr_maxsim_aggregate_device(int device, const float *d_dist, const int64_t *d_keys, int B, int T, int kp,
                               int group_shift, int k_out, double *d_out_score, int64_t *d_out_group, void *stream) {
    if (B < 0 || T < 1 || kp < 1 || k_out < 1 || group_shift < 0 || group_shift > 62)
        return fail(FR_EINVAL, "bad B/T/kp/group_shift/k_out");
    if (static_cast<int64_t>(T) * kp > 1024) return fail(FR_EUNSUP, "T*kp = %d exceeds 1024 candidates per query", T * kp);
    if (B == 0) return FR_OK;
    if (!d_dist || !d_keys || !d_out_score || !d_out_group) return fail(FR_EINVAL, "NULL buffer");
    int rc = check_device(device, nullptr);
    if (rc != FR_OK) return rc;
    DeviceGuard g(device);
    fr::MaxSimArgs ma{d_dist, d_keys, B, T, kp, group_shift, k_out, d_out_score, d_out_group,
                      static_cast<cudaStream_t>(stream)};
    FR_CUDA(fr::launch_maxsim(ma));
    return FR_OK;
}

int fr_maxsim_aggregate(int device, const float *dist, const int64_t *keys, int B, int T, int kp, int group_shift,
                        int k_out, double *out_score, int64_t *out_group) {
    if (B < 0 || T < 1 || kp < 1 || k_out < 1) return fail(FR_EINVAL, "bad B/T/kp/k_out");
    if (B == 0) return FR_OK;
    if (!dist || !keys || !out_score || !out_group) return fail(FR_EINVAL, "NULL buffer");
    const size_t ne = static_cast<size_t>(B) * T * kp;
    return fuse_on_device(device, dist, ne * sizeof(float), keys, ne * sizeof(int64_t), out_score, out_group,
                          static_cast<size_t>(B) * k_out * 8,
                          [&](void *d_d, void *d_k, void *d_sc, void *d_g, cudaStream_t s) {
                              return fr_maxsim_aggregate_device(device, static_cast<const float *>(d_d),
                                                                static_cast<const int64_t *>(d_k), B, T, kp, group_shift, k_out,
                                                                static_cast<double *>(d_sc), static_cast<int64_t *>(d_g), s);
                          });
}

}  // extern "C"
