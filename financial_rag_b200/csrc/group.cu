// fr_group: one collection row-sharded over W B200s (include/fr_index.h, "row-sharded collection").
//
// SURVEY.md 8e / north star: "the corpus is row-sharded across the 8 GPUs of one box; each GPU computes a local top-k,
// then an NCCL all-gather over NVLink feeds a final merge kernel".  A group owns `n_local` of the W shards (all of them
// in the usual single-process case: the reference's Flask server is one process; one of them per process under torchrun)
// and runs one search as
//     queries -> every device   (H2D from one pinned block per device; multi-process: H2D on shard 0 + ncclBroadcast)
//     K1/K2 + K3 per shard      (fr_index_search_partial_device on one stream per device, all devices at once)
//     ncclAllGather             (one grouped call; [packed | keys] = 16 B per (query, result) per shard)
//     K3 in SHARDS mode         (ties -> global insertion order, so the answer equals the one-GPU answer for any W)
// Placement is CYCLIC: global row s (the s-th vector ever inserted) lives on shard s % W at local row s / W.  A collection
// that grows by upserts cannot know its final block boundaries; cyclic placement keeps the shards balanced to within one
// row at every moment and makes the global insertion order of two tied candidates computable from (local row, shard)
// alone (merge_topk.cu).  Upserts of an existing key go to the shard that holds it (overwrite in place).
//
// Exchange modes.  FR_XCHG_NCCL: the grouped all-gather above (every device ends up with every list; the only mode when the
// shards live in several processes).  FR_XCHG_PEER (one process, the default there): no collective at all -- every
// shard's last kernel writes its [packed | keys] lists STRAIGHT INTO the merging GPU's buffer through NVLink peer
// addressing (the output pointers handed to fr_index_search_partial_device are peer pointers), an event per shard tells
// the merging stream when they have landed.  The gather is fused into the producers; what is left of the exchange is W
// event waits.  FR_XCHG_COPY: local lists + cudaMemcpyPeerAsync (kept as the plain reference of the three).
//
// NCCL is bound at run time (dlopen): the process usually has torch's bundled libnccl.so.2 loaded already and two copies
// of NCCL in one process is asking for trouble; fr_nccl_load() names another one.  Groups whose devices are not distinct
// (tests on a one-GPU box) or that ask for it exchange the lists with peer copies instead (FR_XCHG_COPY).
#include <dlfcn.h>

#include <functional>
#include <thread>

#include <algorithm>
#include <condition_variable>
#include <cstring>
#include <mutex>
#include <new>
#include <string>
#include <unordered_map>
#include <vector>

#include "fr_common.cuh"
#include "fr_host.h"
#include "fr_kernels.h"

namespace {

using fr::DevBuf;
using fr::DeviceGuard;
using fr::fail;
using fr::PinBuf;

// ---- NCCL, bound at run time ---------------------------------------------------------------------------------------
typedef struct ncclComm *nccl_comm_t;
typedef struct {
    char internal[128];
} nccl_unique_id;
enum { NCCL_INT64 = 4, NCCL_FLOAT32 = 7 };  // ncclDataType_t values, stable since NCCL 2.0

struct NcclApi {
    void *handle = nullptr;
    int version = 0;
    std::string path;
    int (*GetVersion)(int *) = nullptr;
    int (*GetUniqueId)(nccl_unique_id *) = nullptr;
    int (*CommInitAll)(nccl_comm_t *, int, const int *) = nullptr;
    int (*CommInitRank)(nccl_comm_t *, int, nccl_unique_id, int) = nullptr;
    int (*CommDestroy)(nccl_comm_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    int (*AllGather)(const void *, void *, size_t, int, nccl_comm_t, cudaStream_t) = nullptr;
    int (*Broadcast)(const void *, void *, size_t, int, int, nccl_comm_t, cudaStream_t) = nullptr;
    const char *(*GetErrorString)(int) = nullptr;
};
std::mutex g_nccl_mu;
NcclApi g_nccl;

int nccl_bind(const char *path) {
    std::lock_guard<std::mutex> lk(g_nccl_mu);
    if (g_nccl.handle) return FR_OK;
    void *h = nullptr;
    std::string used;
    if (path && *path) {
        h = dlopen(path, RTLD_NOW | RTLD_GLOBAL);
        used = path;
        if (!h) return fail(FR_EUNSUP, "cannot load NCCL from '%s': %s", path, dlerror());
    } else {
        // the copy the process already holds (torch's), else the system one
        h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);
        used = "libnccl.so.2 (already loaded)";
        if (!h) {
            h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
            used = "libnccl.so.2";
        }
        if (!h) return fail(FR_EUNSUP, "libnccl.so.2 not found (%s); call fr_nccl_load(path) first", dlerror());
    }
    NcclApi a;
    a.handle = h;
    a.path = used;
#define FR_SYM(field, name)                                                        \
    *reinterpret_cast<void **>(&a.field) = dlsym(h, name);                         \
    if (!a.field) return fail(FR_EUNSUP, "%s lacks the symbol %s", used.c_str(), name)
    FR_SYM(GetVersion, "ncclGetVersion");
    FR_SYM(GetUniqueId, "ncclGetUniqueId");
    FR_SYM(CommInitAll, "ncclCommInitAll");
    FR_SYM(CommInitRank, "ncclCommInitRank");
    FR_SYM(CommDestroy, "ncclCommDestroy");
    FR_SYM(GroupStart, "ncclGroupStart");
    FR_SYM(GroupEnd, "ncclGroupEnd");
    FR_SYM(AllGather, "ncclAllGather");
    FR_SYM(Broadcast, "ncclBroadcast");
    FR_SYM(GetErrorString, "ncclGetErrorString");
#undef FR_SYM
    a.GetVersion(&a.version);
    g_nccl = a;
    return FR_OK;
}

#define FR_NCCL(expr)                                                                                       \
    do {                                                                                                    \
        int _r = (expr);                                                                                    \
        if (_r != 0)                                                                                        \
            return fail(FR_ECUDA, "%s failed: %s (%s:%d)", #expr, g_nccl.GetErrorString(_r), __FILE__, __LINE__); \
    } while (0)

constexpr int64_t SHARD_ROW_LIMIT = 0xfffffff0ll;

}  // namespace

// One enqueueing thread per local shard beyond the first: a local search is ~11 kernel launches, and eight of them issued
// back to back from one thread cost a batch-1 search over 8 GPUs 0.4 ms of host time (1.83 ms against 1.44 ms with one
// process per GPU, profiles/r02_bench_cfg4_n8_single_process.json).
struct ShardWorker {
    std::thread th;
    std::mutex m;
    std::condition_variable cv;
    std::function<void()> job;
    bool has_job = false, done = true, quit = false;
    void loop() {
        std::unique_lock<std::mutex> lk(m);
        for (;;) {
            cv.wait(lk, [&] { return has_job || quit; });
            if (quit) return;
            std::function<void()> j = std::move(job);
            has_job = false;
            lk.unlock();
            j();
            lk.lock();
            done = true;
            cv.notify_all();
        }
    }
    void submit(std::function<void()> j) {
        std::lock_guard<std::mutex> lk(m);
        job = std::move(j);
        has_job = true;
        done = false;
        cv.notify_all();
    }
    void wait() {
        std::unique_lock<std::mutex> lk(m);
        cv.wait(lk, [&] { return done; });
    }
    void stop() {
        {
            std::lock_guard<std::mutex> lk(m);
            quit = true;
        }
        cv.notify_all();
        if (th.joinable()) th.join();
    }
};

struct fr_group {
    int dim = 0, metric = FR_COSINE, dtype = FR_BF16;
    int W = 1, first = 0, nl = 1;  // shards in the world, first local shard, local shards
    int exchange = FR_XCHG_COPY;
    std::vector<int> dev;
    std::vector<fr_index *> shard;
    std::vector<cudaStream_t> stream;   // one internal stream per local shard (host entry points)
    std::vector<nccl_comm_t> comm;      // one communicator rank per local shard (FR_XCHG_NCCL)
    std::vector<cudaEvent_t> ev_use;    // last use of the shard's exchange scratch (orders caller streams)
    std::vector<cudaEvent_t> ev_local;  // FR_XCHG_COPY: the shard's local lists are complete
    std::vector<cudaEvent_t> ev_done;   // FR_XCHG_COPY: device j has finished reading everybody's lists
    std::vector<DevBuf> q, send, recv, out_dist, out_keys;
    // host searches: `mu` is held only while a call is enqueued; each call owns one pinned slot and waits for its own event
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
    int64_t rows = 0;       // global rows, deleted ones included
    int64_t n_deleted = 0;  // global
    std::unordered_map<int64_t, int64_t> keymap;  // key -> global row (live rows only)
    bool keymap_valid = true;
    int64_t n_searches = 0;
    bool enqueue_threads = true;
    std::vector<ShardWorker *> workers;  // [nl], entry 0 unused (the calling thread serves shard 0); created on first use

    bool all_local() const { return nl == W; }
    size_t row_bytes() const { return static_cast<size_t>(dim) * (dtype == FR_BF16 ? 2 : 4); }
    // global rows [0, total) that live on world shard s
    int64_t rows_of(int s, int64_t total) const { return total > s ? (total - s + W - 1) / W : 0; }
};

namespace {

int group_rebuild_keymap(fr_group *g) {
    if (g->keymap_valid) return FR_OK;
    if (!g->all_local())
        return fail(FR_EUNSUP, "a bulk-loaded group whose shards live in several processes has no key map: "
                               "upsert / delete / lookup need every shard in one process");
    g->keymap.clear();
    g->keymap.reserve(static_cast<size_t>(g->rows) * 2);
    std::vector<int64_t> hk;
    for (int j = 0; j < g->nl; ++j) {
        const int64_t n = g->rows_of(j, g->rows);
        hk.resize(static_cast<size_t>(n));
        if (n == 0) continue;
        int rc = fr_index_export_raw(g->shard[j], 0, n, nullptr, hk.data());
        if (rc != FR_OK) return rc;
        for (int64_t r = 0; r < n; ++r)
            if (hk[r] != fr::KEY_TOMBSTONE) g->keymap[hk[r]] = r * g->W + j;
    }
    g->keymap_valid = true;
    return FR_OK;
}

int check_group_search_args(fr_group *g, int B, int k) {
    if (!g) return fail(FR_EINVAL, "group is NULL");
    if (B < 0) return fail(FR_EINVAL, "B = %d is negative", B);
    if (k < 1) return fail(FR_EINVAL, "k = %d must be >= 1", k);
    if (k > FR_MAX_K) return fail(FR_EUNSUP, "k = %d exceeds FR_MAX_K = %d", k, FR_MAX_K);
    return FR_OK;
}

// The device part of a search: queries already on every local device.  streams[j] is the stream of local shard j.
// Where d_out_keys[j] is non-NULL the merged result is written on device j.
int group_search_enqueue(fr_group *g, const float *const *d_queries, int B, int k, float *const *d_out_dist,
                         int64_t *const *d_out_keys, const cudaStream_t *streams) {
    const size_t m = static_cast<size_t>(B) * k;  // entries per shard and field
    for (int j = 0; j < g->nl; ++j) {
        DeviceGuard dg(g->dev[j]);
        FR_CUDA(g->send[j].need(2 * m * sizeof(int64_t)));
        const bool gathers = g->exchange == FR_XCHG_NCCL || d_out_keys[j] != nullptr;
        if (gathers) FR_CUDA(g->recv[j].need(static_cast<size_t>(g->W) * 2 * m * sizeof(int64_t)));
        FR_CUDA(cudaStreamWaitEvent(streams[j], g->ev_use[j], 0));
        if (g->exchange != FR_XCHG_NCCL)  // nobody may still be reading the lists this call is about to overwrite
            for (int o = 0; o < g->nl; ++o) FR_CUDA(cudaStreamWaitEvent(streams[j], g->ev_done[o], 0));
    }
    int root = -1;  // FR_XCHG_PEER: the device whose buffer every shard writes into (the first one that wants the result)
    if (g->exchange == FR_XCHG_PEER)
        for (int o = 0; o < g->nl && root < 0; ++o)
            if (d_out_keys[o] != nullptr) root = o;
    // the local searches: shard 0 from this thread, the others from their own enqueueing threads, all at once
    const bool threaded = g->nl > 1 && g->enqueue_threads;
    if (threaded && g->workers.empty()) {
        g->workers.assign(static_cast<size_t>(g->nl), nullptr);
        for (int j = 1; j < g->nl; ++j) {
            g->workers[static_cast<size_t>(j)] = new ShardWorker();
            ShardWorker *w = g->workers[static_cast<size_t>(j)];
            w->th = std::thread([w] { w->loop(); });
        }
    }
    std::vector<int> rcs(static_cast<size_t>(g->nl), FR_OK);
    std::vector<std::string> errs(static_cast<size_t>(g->nl));
    auto local = [&, root, m](int j) {
        uint64_t *packed = static_cast<uint64_t *>(g->send[j].p);
        int64_t *keys = static_cast<int64_t *>(g->send[j].p) + m;
        if (root >= 0) {  // peer pointers: the shard's last kernel stores into the merging GPU's memory
            packed = static_cast<uint64_t *>(g->recv[root].p) + static_cast<size_t>(j) * 2 * m;
            keys = reinterpret_cast<int64_t *>(packed) + m;
        }
        rcs[static_cast<size_t>(j)] = fr_index_search_partial_device(g->shard[j], d_queries[j], B, k, packed, keys, streams[j]);
        if (rcs[static_cast<size_t>(j)] != FR_OK) errs[static_cast<size_t>(j)] = fr_last_error();  // (thread-local: carry it over)
    };
    if (threaded) {
        for (int j = 1; j < g->nl; ++j) g->workers[static_cast<size_t>(j)]->submit([&local, j] { local(j); });
        local(0);
        for (int j = 1; j < g->nl; ++j) g->workers[static_cast<size_t>(j)]->wait();
    } else {
        for (int j = 0; j < g->nl; ++j) local(j);
    }
    for (int j = 0; j < g->nl; ++j)
        if (rcs[static_cast<size_t>(j)] != FR_OK)
            return fail(rcs[static_cast<size_t>(j)], "shard %d: %s", g->first + j, errs[static_cast<size_t>(j)].c_str());
    if (g->exchange == FR_XCHG_NCCL) {
        FR_NCCL(g_nccl.GroupStart());
        for (int j = 0; j < g->nl; ++j) {
            int r = g_nccl.AllGather(g->send[j].p, g->recv[j].p, 2 * m, NCCL_INT64, g->comm[j], streams[j]);
            if (r != 0) {
                g_nccl.GroupEnd();
                return fail(FR_ECUDA, "ncclAllGather failed: %s", g_nccl.GetErrorString(r));
            }
        }
        FR_NCCL(g_nccl.GroupEnd());
    } else {
        for (int j = 0; j < g->nl; ++j) {
            DeviceGuard dg(g->dev[j]);
            FR_CUDA(cudaEventRecord(g->ev_local[j], streams[j]));
        }
        for (int o = 0; o < g->nl; ++o) {
            if (d_out_keys[o] == nullptr) continue;
            DeviceGuard dg(g->dev[o]);
            for (int j = 0; j < g->nl; ++j)
                if (j != o) FR_CUDA(cudaStreamWaitEvent(streams[o], g->ev_local[j], 0));
            if (root == o) continue;  // the lists are already here
            if (root >= 0) {          // a second consumer: one copy of the gathered block from the first
                FR_CUDA(cudaMemcpyPeerAsync(g->recv[o].p, g->dev[o], g->recv[root].p, g->dev[root],
                                            static_cast<size_t>(g->W) * 2 * m * sizeof(int64_t), streams[o]));
                continue;
            }
            for (int j = 0; j < g->nl; ++j) {
                uint8_t *dst = static_cast<uint8_t *>(g->recv[o].p) + static_cast<size_t>(j) * 2 * m * sizeof(int64_t);
                FR_CUDA(cudaMemcpyPeerAsync(dst, g->dev[o], g->send[j].p, g->dev[j], 2 * m * sizeof(int64_t), streams[o]));
            }
        }
    }
    for (int o = 0; o < g->nl; ++o) {
        DeviceGuard dg(g->dev[o]);
        if (d_out_keys[o] != nullptr) {
            fr::MergeArgs ma{};
            ma.packed = static_cast<const uint64_t *>(g->recv[o].p);
            ma.P = g->W;
            ma.shard_stride = static_cast<int64_t>(2 * m);
            ma.B = B;
            ma.k = k;
            ma.shards = true;
            ma.shard_keys = static_cast<const int64_t *>(g->recv[o].p) + m;
            ma.cyclic_world = g->W;
            ma.l2 = g->metric == FR_L2;
            ma.out_dist = d_out_dist[o];
            ma.out_keys = d_out_keys[o];
            ma.stream = streams[o];
            FR_CUDA(fr::launch_merge_topk(ma));
            if (g->exchange != FR_XCHG_NCCL) FR_CUDA(cudaEventRecord(g->ev_done[o], streams[o]));
        }
        FR_CUDA(cudaEventRecord(g->ev_use[o], streams[o]));
    }
    g->n_searches += 1;
    return FR_OK;
}

}  // namespace

extern "C" {

int fr_nccl_load(const char *path) { return nccl_bind(path); }

int fr_nccl_version(int *out) {
    if (!out) return fail(FR_EINVAL, "out is NULL");
    int rc = nccl_bind(nullptr);
    if (rc != FR_OK) return rc;
    *out = g_nccl.version;
    return FR_OK;
}

int fr_nccl_unique_id(void *out, int nbytes) {
    if (!out || nbytes < static_cast<int>(sizeof(nccl_unique_id)))
        return fail(FR_EINVAL, "the unique id needs a buffer of %d bytes", (int)sizeof(nccl_unique_id));
    int rc = nccl_bind(nullptr);
    if (rc != FR_OK) return rc;
    nccl_unique_id id;
    FR_NCCL(g_nccl.GetUniqueId(&id));
    std::memcpy(out, &id, sizeof(id));
    return FR_OK;
}

int fr_group_destroy(fr_group *g) {
    if (!g) return FR_OK;
    for (ShardWorker *w : g->workers)
        if (w) {
            w->stop();
            delete w;
        }
    g->workers.clear();
    for (int j = 0; j < static_cast<int>(g->shard.size()); ++j) {
        DeviceGuard dg(g->dev[j]);
        cudaDeviceSynchronize();
    }
    for (nccl_comm_t c : g->comm)
        if (c && g_nccl.CommDestroy) g_nccl.CommDestroy(c);
    for (int j = 0; j < static_cast<int>(g->dev.size()); ++j) {
        DeviceGuard dg(g->dev[j]);
        if (j < static_cast<int>(g->shard.size()) && g->shard[j]) fr_index_destroy(g->shard[j]);
        auto rel = [j](std::vector<DevBuf> &v) {
            if (j < static_cast<int>(v.size())) v[j].release();
        };
        rel(g->q);
        rel(g->send);
        rel(g->recv);
        rel(g->out_dist);
        rel(g->out_keys);
        for (auto *evs : {&g->ev_use, &g->ev_local, &g->ev_done})
            if (j < static_cast<int>(evs->size()) && (*evs)[j]) cudaEventDestroy((*evs)[j]);
        if (j < static_cast<int>(g->stream.size()) && g->stream[j]) cudaStreamDestroy(g->stream[j]);
    }
    if (!g->dev.empty()) {
        DeviceGuard dg(g->dev[0]);
        for (auto &sl : g->slots) {
            sl.pin.release();
            if (sl.done) cudaEventDestroy(sl.done);
        }
    }
    delete g;
    return FR_OK;
}

int fr_group_create(int dim, int metric, int dtype, const int *devices, int n_local, int world_shards, int first_shard,
                    const void *nccl_id, int exchange, int64_t reserve_rows_per_shard, fr_group **out) {
    if (!out) return fail(FR_EINVAL, "out is NULL");
    *out = nullptr;
    if (!devices || n_local < 1) return fail(FR_EINVAL, "a group needs at least one local device");
    if (world_shards <= 0) world_shards = n_local;
    if (world_shards > 64) return fail(FR_EUNSUP, "at most 64 shards per group");
    if (first_shard < 0 || first_shard + n_local > world_shards)
        return fail(FR_EINVAL, "local shards [%d, %d) do not fit a world of %d", first_shard, first_shard + n_local,
                    world_shards);
    if (exchange < FR_XCHG_AUTO || exchange > FR_XCHG_PEER) return fail(FR_EINVAL, "unknown exchange mode %d", exchange);
    const bool all_local = n_local == world_shards;
    if (!all_local && !nccl_id)
        return fail(FR_EINVAL, "a group spanning several processes needs the NCCL unique id of its rank 0 (fr_nccl_unique_id)");
    bool distinct = true;
    for (int a = 0; a < n_local; ++a)
        for (int b = a + 1; b < n_local; ++b)
            if (devices[a] == devices[b]) distinct = false;
    // can every local device address every other one's memory?  (the same device twice: trivially)
    bool peers_ok = true;
    for (int a = 0; a < n_local && peers_ok; ++a)
        for (int b = 0; b < n_local; ++b) {
            if (devices[a] == devices[b]) continue;
            int can = 0;
            if (cudaDeviceCanAccessPeer(&can, devices[a], devices[b]) != cudaSuccess || !can) {
                cudaGetLastError();
                peers_ok = false;
                break;
            }
        }
    if (exchange == FR_XCHG_AUTO) {
        if (!all_local) exchange = FR_XCHG_NCCL;
        else if (peers_ok) exchange = FR_XCHG_PEER;
        else exchange = distinct ? FR_XCHG_NCCL : FR_XCHG_COPY;
    }
    if ((exchange == FR_XCHG_COPY || exchange == FR_XCHG_PEER) && !all_local)
        return fail(FR_EINVAL, "FR_XCHG_COPY / FR_XCHG_PEER need every shard in this process");
    if (exchange == FR_XCHG_PEER && !peers_ok) return fail(FR_EUNSUP, "FR_XCHG_PEER needs peer access between all devices of the group");
    if (exchange == FR_XCHG_NCCL && !distinct)
        return fail(FR_EINVAL, "NCCL needs distinct devices (one communicator rank per GPU); use FR_XCHG_COPY");
    if (exchange == FR_XCHG_NCCL) {
        int rc = nccl_bind(nullptr);
        if (rc != FR_OK) return rc;
    }
    fr_group *g = new (std::nothrow) fr_group();
    if (!g) return fail(FR_ENOMEM, "host allocation failed");
    g->dim = dim;
    g->metric = metric;
    g->dtype = dtype;
    g->W = world_shards;
    g->first = first_shard;
    g->nl = n_local;
    g->exchange = exchange;
    g->dev.assign(devices, devices + n_local);
    g->q.resize(n_local);
    g->send.resize(n_local);
    g->recv.resize(n_local);
    g->out_dist.resize(n_local);
    g->out_keys.resize(n_local);
    g->stream.assign(n_local, nullptr);
    g->ev_use.assign(n_local, nullptr);
    g->ev_local.assign(n_local, nullptr);
    g->ev_done.assign(n_local, nullptr);
    for (int j = 0; j < n_local; ++j) {
        fr_index *ix = nullptr;
        int rc = fr_index_create(dim, metric, dtype, devices[j], reserve_rows_per_shard, &ix);
        if (rc != FR_OK) {
            const std::string msg = fr_last_error();
            fr_group_destroy(g);
            return fail(rc, "shard %d on device %d: %s", first_shard + j, devices[j], msg.c_str());
        }
        g->shard.push_back(ix);
        DeviceGuard dg(devices[j]);
        cudaError_t e = cudaStreamCreateWithFlags(&g->stream[j], cudaStreamNonBlocking);
        for (cudaEvent_t *ev : {&g->ev_use[j], &g->ev_local[j], &g->ev_done[j]}) {
            if (e == cudaSuccess) e = cudaEventCreateWithFlags(ev, cudaEventDisableTiming);
            if (e == cudaSuccess) e = cudaEventRecord(*ev, g->stream[j]);
        }
        if (e != cudaSuccess) {
            fr_group_destroy(g);
            return fail(FR_ECUDA, "stream/event creation on device %d failed: %s", devices[j], cudaGetErrorString(e));
        }
        // peer access makes the list exchange (and NCCL's own transport) go over NVLink directly
        for (int b = 0; b < n_local; ++b) {
            if (devices[b] == devices[j]) continue;
            int can = 0;
            if (cudaDeviceCanAccessPeer(&can, devices[j], devices[b]) == cudaSuccess && can) {
                cudaError_t pe = cudaDeviceEnablePeerAccess(devices[b], 0);
                if (pe != cudaSuccess) cudaGetLastError();  // already enabled is fine
            }
        }
    }
    if (exchange == FR_XCHG_NCCL) {
        g->comm.assign(n_local, nullptr);
        int r = 0;
        if (all_local) {
            r = g_nccl.CommInitAll(g->comm.data(), n_local, devices);
        } else {
            nccl_unique_id id;
            std::memcpy(&id, nccl_id, sizeof(id));
            r = g_nccl.GroupStart();
            for (int j = 0; j < n_local && r == 0; ++j) {
                DeviceGuard dg(devices[j]);
                r = g_nccl.CommInitRank(&g->comm[j], world_shards, id, first_shard + j);
            }
            const int r2 = g_nccl.GroupEnd();
            if (r == 0) r = r2;
        }
        if (r != 0) {
            const std::string msg = g_nccl.GetErrorString(r);
            fr_group_destroy(g);
            return fail(FR_ECUDA, "NCCL communicator creation failed: %s", msg.c_str());
        }
    }
    *out = g;
    return FR_OK;
}

int fr_group_info(fr_group *g, const char *name, int64_t *out) {
    if (!g || !name || !out) return fail(FR_EINVAL, "NULL argument");
    std::lock_guard<std::mutex> lk(g->mu);
    const std::string n(name);
    if (n == "world_shards") *out = g->W;
    else if (n == "local_shards") *out = g->nl;
    else if (n == "first_shard") *out = g->first;
    else if (n == "exchange") *out = g->exchange;
    else if (n == "rows") *out = g->rows;
    else if (n == "count") *out = g->rows - g->n_deleted;
    else if (n == "searches") *out = g->n_searches;
    else if (n == "nccl_version") *out = g->exchange == FR_XCHG_NCCL ? g_nccl.version : 0;
    else return fail(FR_EINVAL, "unknown group info '%s'", name);
    return FR_OK;
}

int fr_group_shard(fr_group *g, int local_shard, fr_index **out) {
    if (!g || !out) return fail(FR_EINVAL, "NULL argument");
    if (local_shard < 0 || local_shard >= g->nl) return fail(FR_EINVAL, "local shard %d out of range", local_shard);
    *out = g->shard[local_shard];
    return FR_OK;
}

int fr_group_set_option(fr_group *g, const char *name, int64_t value) {
    if (!g || !name) return fail(FR_EINVAL, "NULL argument");
    if (std::strcmp(name, "enqueue_threads") == 0) {  // 0: this thread enqueues every local search itself (for A/B timing)
        std::lock_guard<std::mutex> lk(g->mu);
        g->enqueue_threads = value != 0;
        return FR_OK;
    }
    for (fr_index *ix : g->shard) {
        int rc = fr_index_set_option(ix, name, value);
        if (rc != FR_OK) return rc;
    }
    return FR_OK;
}

int fr_group_reserve(fr_group *g, int64_t total_rows) {
    if (!g) return fail(FR_EINVAL, "group is NULL");
    if (total_rows < 0) return fail(FR_EINVAL, "rows is negative");
    for (int j = 0; j < g->nl; ++j) {
        int rc = fr_index_reserve(g->shard[j], g->rows_of(g->first + j, total_rows));
        if (rc != FR_OK) return rc;
    }
    return FR_OK;
}

int fr_group_adopt_rows(fr_group *g, int64_t total_rows) {
    if (!g) return fail(FR_EINVAL, "group is NULL");
    std::lock_guard<std::mutex> lk(g->mu);
    int64_t deleted = 0;
    for (int j = 0; j < g->nl; ++j) {
        int64_t have = 0, live = 0;
        int rc = fr_index_rows(g->shard[j], &have);
        if (rc == FR_OK) rc = fr_index_count(g->shard[j], &live);
        if (rc != FR_OK) return rc;
        const int64_t want = g->rows_of(g->first + j, total_rows);
        if (have != want)
            return fail(FR_EINVAL, "shard %d holds %lld rows; cyclic placement of %lld rows over %d shards gives it %lld",
                        g->first + j, (long long)have, (long long)total_rows, g->W, (long long)want);
        deleted += have - live;
    }
    if (g->rows_of(0, total_rows) > SHARD_ROW_LIMIT / g->W)
        return fail(FR_EUNSUP, "a shard of a %d-way group holds at most %lld rows", g->W, (long long)(SHARD_ROW_LIMIT / g->W));
    g->rows = total_rows;
    g->n_deleted = deleted;
    g->keymap.clear();
    g->keymap_valid = total_rows == 0;
    return FR_OK;
}

int fr_group_count(fr_group *g, int64_t *out) { return fr_group_info(g, "count", out); }
int fr_group_rows(fr_group *g, int64_t *out) { return fr_group_info(g, "rows", out); }

int fr_group_upsert(fr_group *g, const float *vecs, const int64_t *keys, int64_t n) {
    if (!g) return fail(FR_EINVAL, "group is NULL");
    if (n < 0) return fail(FR_EINVAL, "n is negative");
    if (n == 0) return FR_OK;
    if (!vecs || !keys) return fail(FR_EINVAL, "NULL buffer");
    std::lock_guard<std::mutex> lk(g->mu);
    int rc = group_rebuild_keymap(g);
    if (rc != FR_OK) return rc;
    for (int64_t i = 0; i < n; ++i)
        if (keys[i] == fr::KEY_TOMBSTONE || keys[i] == FR_KEY_NONE)
            return fail(FR_EINVAL, "key %lld is reserved", (long long)keys[i]);
    // existing key -> the shard that holds it; new key -> global row `rows`, `rows + 1`, ... (shard = row % W)
    int64_t new_rows = g->rows;
    std::vector<std::vector<int64_t>> take(static_cast<size_t>(g->nl));
    std::vector<int64_t> added;
    for (int64_t i = 0; i < n; ++i) {
        int64_t row;
        auto it = g->keymap.find(keys[i]);
        if (it != g->keymap.end()) {
            row = it->second;
        } else {
            row = new_rows++;
            g->keymap.emplace(keys[i], row);
            added.push_back(keys[i]);
        }
        const int s = static_cast<int>(row % g->W) - g->first;
        if (s >= 0 && s < g->nl) take[static_cast<size_t>(s)].push_back(i);
    }
    auto undo = [&]() {
        for (int64_t key : added) g->keymap.erase(key);
    };
    if (g->rows_of(0, new_rows) > SHARD_ROW_LIMIT / g->W) {
        undo();
        return fail(FR_EUNSUP, "a shard of a %d-way group holds at most %lld rows", g->W, (long long)(SHARD_ROW_LIMIT / g->W));
    }
    std::vector<float> sv;
    std::vector<int64_t> sk;
    for (int j = 0; j < g->nl; ++j) {
        const std::vector<int64_t> &idx = take[static_cast<size_t>(j)];
        if (idx.empty()) continue;
        sv.resize(idx.size() * static_cast<size_t>(g->dim));
        sk.resize(idx.size());
        for (size_t t = 0; t < idx.size(); ++t) {
            std::memcpy(sv.data() + t * g->dim, vecs + idx[t] * g->dim, static_cast<size_t>(g->dim) * sizeof(float));
            sk[t] = keys[idx[t]];
        }
        rc = fr_index_upsert(g->shard[j], sv.data(), sk.data(), static_cast<int64_t>(idx.size()));
        if (rc != FR_OK) {
            g->keymap_valid = false;  // some shards took their rows, this one did not: rebuild from the devices
            return rc;
        }
    }
    g->rows = new_rows;
    return FR_OK;
}

int fr_group_delete(fr_group *g, const int64_t *keys, int64_t n, int64_t *out_deleted) {
    if (!g) return fail(FR_EINVAL, "group is NULL");
    if (out_deleted) *out_deleted = 0;
    if (n < 0) return fail(FR_EINVAL, "n is negative");
    if (n == 0) return FR_OK;
    if (!keys) return fail(FR_EINVAL, "NULL buffer");
    std::lock_guard<std::mutex> lk(g->mu);
    int rc = group_rebuild_keymap(g);
    if (rc != FR_OK) return rc;
    std::vector<std::vector<int64_t>> take(static_cast<size_t>(g->nl));
    int64_t found = 0;
    for (int64_t i = 0; i < n; ++i) {
        auto it = g->keymap.find(keys[i]);
        if (it == g->keymap.end()) continue;
        const int s = static_cast<int>(it->second % g->W) - g->first;
        if (s >= 0 && s < g->nl) take[static_cast<size_t>(s)].push_back(keys[i]);
        g->keymap.erase(it);
        ++found;
    }
    for (int j = 0; j < g->nl; ++j) {
        if (take[static_cast<size_t>(j)].empty()) continue;
        int64_t d = 0;
        rc = fr_index_delete(g->shard[j], take[static_cast<size_t>(j)].data(), static_cast<int64_t>(take[static_cast<size_t>(j)].size()), &d);
        if (rc != FR_OK) {
            g->keymap_valid = false;
            return rc;
        }
    }
    g->n_deleted += found;
    if (out_deleted) *out_deleted = found;
    return FR_OK;
}

// rows [first_row, first_row + n) in global insertion order; which of them live on local shard j, and where
static void shard_span(const fr_group *g, int j, int64_t first_row, int64_t n, int64_t *g0, int64_t *cnt, int64_t *l0) {
    const int s = g->first + j;
    int64_t r0 = first_row + ((s - first_row % g->W) % g->W + g->W) % g->W;  // first global row >= first_row on shard s
    *g0 = r0;
    *cnt = r0 < first_row + n ? (first_row + n - r0 + g->W - 1) / g->W : 0;
    *l0 = r0 / g->W;
}

static int group_read_rows(fr_group *g, int64_t first_row, int64_t n, void *out_rows, size_t elem_row_bytes, bool widen,
                           int64_t *out_keys) {
    if (!g) return fail(FR_EINVAL, "group is NULL");
    std::lock_guard<std::mutex> lk(g->mu);
    if (!g->all_local()) return fail(FR_EUNSUP, "reading rows in global order needs every shard in one process");
    if (first_row < 0 || n < 0 || first_row + n > g->rows)
        return fail(FR_EINVAL, "rows [%lld, %lld) out of range (group holds %lld)", (long long)first_row,
                    (long long)(first_row + n), (long long)g->rows);
    if (n == 0) return FR_OK;
    std::vector<uint8_t> tmp;
    std::vector<int64_t> tk;
    for (int j = 0; j < g->nl; ++j) {
        int64_t g0, cnt, l0;
        shard_span(g, j, first_row, n, &g0, &cnt, &l0);
        if (cnt == 0) continue;
        if (out_rows) tmp.resize(static_cast<size_t>(cnt) * elem_row_bytes);
        if (out_keys) tk.resize(static_cast<size_t>(cnt));
        int rc = widen ? fr_index_get_rows(g->shard[j], l0, cnt, out_rows ? reinterpret_cast<float *>(tmp.data()) : nullptr,
                                           out_keys ? tk.data() : nullptr)
                       : fr_index_export_raw(g->shard[j], l0, cnt, out_rows ? tmp.data() : nullptr,
                                             out_keys ? tk.data() : nullptr);
        if (rc != FR_OK) return rc;
        for (int64_t t = 0; t < cnt; ++t) {
            const int64_t at = g0 + t * g->W - first_row;
            if (out_rows)
                std::memcpy(static_cast<uint8_t *>(out_rows) + static_cast<size_t>(at) * elem_row_bytes,
                            tmp.data() + static_cast<size_t>(t) * elem_row_bytes, elem_row_bytes);
            if (out_keys) out_keys[at] = tk[static_cast<size_t>(t)];
        }
    }
    return FR_OK;
}

int fr_group_export_raw(fr_group *g, int64_t first_row, int64_t n, void *out_rows, int64_t *out_keys) {
    return group_read_rows(g, first_row, n, out_rows, g ? g->row_bytes() : 0, false, out_keys);
}

int fr_group_get_rows(fr_group *g, int64_t first_row, int64_t n, float *out_vecs, int64_t *out_keys) {
    return group_read_rows(g, first_row, n, out_vecs, g ? static_cast<size_t>(g->dim) * sizeof(float) : 0, true, out_keys);
}

int fr_group_import_raw(fr_group *g, const void *rows, const int64_t *keys, int64_t n) {
    if (!g) return fail(FR_EINVAL, "group is NULL");
    if (n < 0) return fail(FR_EINVAL, "n is negative");
    if (n == 0) return FR_OK;
    if (!rows || !keys) return fail(FR_EINVAL, "NULL buffer");
    std::lock_guard<std::mutex> lk(g->mu);
    if (g->rows_of(0, g->rows + n) > SHARD_ROW_LIMIT / g->W)
        return fail(FR_EUNSUP, "a shard of a %d-way group holds at most %lld rows", g->W, (long long)(SHARD_ROW_LIMIT / g->W));
    const size_t rb = g->row_bytes();
    std::vector<uint8_t> tmp;
    std::vector<int64_t> tk;
    for (int j = 0; j < g->nl; ++j) {
        int64_t g0, cnt, l0;
        shard_span(g, j, g->rows, n, &g0, &cnt, &l0);
        if (cnt == 0) continue;
        tmp.resize(static_cast<size_t>(cnt) * rb);
        tk.resize(static_cast<size_t>(cnt));
        for (int64_t t = 0; t < cnt; ++t) {
            const int64_t at = g0 + t * g->W - g->rows;
            std::memcpy(tmp.data() + static_cast<size_t>(t) * rb, static_cast<const uint8_t *>(rows) + static_cast<size_t>(at) * rb, rb);
            tk[static_cast<size_t>(t)] = keys[at];
        }
        int rc = fr_index_import_raw(g->shard[j], tmp.data(), tk.data(), cnt);
        if (rc != FR_OK) {
            g->keymap_valid = false;
            return rc;
        }
    }
    for (int64_t i = 0; i < n; ++i)
        if (keys[i] == fr::KEY_TOMBSTONE) g->n_deleted += 1;
    g->rows += n;
    g->keymap.clear();
    g->keymap_valid = false;
    return FR_OK;
}

int fr_group_lookup_rows(fr_group *g, const int64_t *keys, int64_t n, int64_t *out_rows) {
    if (!g) return fail(FR_EINVAL, "group is NULL");
    if (n < 0) return fail(FR_EINVAL, "n is negative");
    if (n == 0) return FR_OK;
    if (!keys || !out_rows) return fail(FR_EINVAL, "NULL buffer");
    std::lock_guard<std::mutex> lk(g->mu);
    int rc = group_rebuild_keymap(g);
    if (rc != FR_OK) return rc;
    for (int64_t i = 0; i < n; ++i) {
        auto it = g->keymap.find(keys[i]);
        out_rows[i] = it == g->keymap.end() ? -1 : it->second;
    }
    return FR_OK;
}

int fr_group_search_device(fr_group *g, const float *const *d_queries, int B, int k, float *const *d_out_dist,
                           int64_t *const *d_out_keys, void *const *streams) {
    int rc = check_group_search_args(g, B, k);
    if (rc != FR_OK || B == 0) return rc;
    if (!d_queries || !d_out_dist || !d_out_keys) return fail(FR_EINVAL, "NULL pointer table");
    std::vector<cudaStream_t> st(static_cast<size_t>(g->nl));
    for (int j = 0; j < g->nl; ++j) {
        if (!d_queries[j]) return fail(FR_EINVAL, "queries missing on local shard %d", j);
        if ((d_out_dist[j] == nullptr) != (d_out_keys[j] == nullptr)) return fail(FR_EINVAL, "give both outputs of a device or neither");
        st[static_cast<size_t>(j)] = streams ? static_cast<cudaStream_t>(streams[j]) : g->stream[j];
    }
    std::lock_guard<std::mutex> lk(g->mu);
    return group_search_enqueue(g, d_queries, B, k, d_out_dist, d_out_keys, st.data());
}

int fr_group_search(fr_group *g, const float *queries, int B, int k, float *out_dist, int64_t *out_keys) {
    int rc = check_group_search_args(g, B, k);
    if (rc != FR_OK || B == 0) return rc;
    const bool has_root = g->first == 0;  // the process that holds shard 0 feeds the queries
    if (has_root && !queries) return fail(FR_EINVAL, "NULL queries");
    if (!out_dist || !out_keys) return fail(FR_EINVAL, "NULL buffer");
    // a pinned staging slot of this call alone (concurrent callers overlap their copies and waits with each other's GPU work)
    int sid = -1;
    {
        std::unique_lock<std::mutex> lk(g->slot_mu);
        g->slot_cv.wait(lk, [&] {
            for (auto &sl : g->slots)
                if (!sl.busy) return true;
            return false;
        });
        for (int i = 0; i < fr_group::N_SLOTS; ++i)
            if (!g->slots[i].busy) {
                g->slots[i].busy = true;
                sid = i;
                break;
            }
    }
    struct Release {
        fr_group *g;
        int sid;
        ~Release() {
            {
                std::lock_guard<std::mutex> lk(g->slot_mu);
                g->slots[sid].busy = false;
            }
            g->slot_cv.notify_one();
        }
    } release{g, sid};
    fr_group::HostSlot &slot = g->slots[sid];
    const size_t qb = static_cast<size_t>(B) * g->dim * sizeof(float);
    const size_t db = static_cast<size_t>(B) * k * sizeof(float), kb = static_cast<size_t>(B) * k * sizeof(int64_t);
    const size_t db_al = (db + 15) & ~static_cast<size_t>(15), kb_al = (kb + 15) & ~static_cast<size_t>(15);
    if (!slot.done || slot.pin.bytes < qb + db_al + kb_al) {
        // rare: growing the block frees and allocates pinned memory (device-wide synchronisation) -- not beside another
        // caller's enqueue (see fr_index_search)
        std::lock_guard<std::mutex> lk(g->mu);
        DeviceGuard dg(g->dev[0]);
        if (!slot.done) FR_CUDA(cudaEventCreateWithFlags(&slot.done, cudaEventDisableTiming));
        FR_CUDA(slot.pin.need(qb + db_al + kb_al));
    }
    uint8_t *pin_keys = static_cast<uint8_t *>(slot.pin.p), *pin_dist = pin_keys + kb_al, *pin_q = pin_dist + db_al;
    if (has_root) std::memcpy(pin_q, queries, qb);
    {
        std::lock_guard<std::mutex> lk(g->mu);
        std::vector<const float *> dq(static_cast<size_t>(g->nl));
        std::vector<float *> od(static_cast<size_t>(g->nl), nullptr);
        std::vector<int64_t *> ok(static_cast<size_t>(g->nl), nullptr);
        for (int j = 0; j < g->nl; ++j) {
            DeviceGuard dg(g->dev[j]);
            FR_CUDA(g->q[j].need(qb));
            dq[static_cast<size_t>(j)] = static_cast<const float *>(g->q[j].p);
        }
        {
            DeviceGuard dg(g->dev[0]);
            FR_CUDA(g->out_dist[0].need(db));
            FR_CUDA(g->out_keys[0].need(kb));
            od[0] = static_cast<float *>(g->out_dist[0].p);
            ok[0] = static_cast<int64_t *>(g->out_keys[0].p);
        }
        if (g->all_local()) {
            // every device pulls the block over its own PCIe link at once
            for (int j = 0; j < g->nl; ++j) {
                DeviceGuard dg(g->dev[j]);
                FR_CUDA(cudaStreamWaitEvent(g->stream[j], g->ev_use[j], 0));
                FR_CUDA(cudaMemcpyAsync(g->q[j].p, pin_q, qb, cudaMemcpyHostToDevice, g->stream[j]));
            }
        } else {
            for (int j = 0; j < g->nl; ++j) {
                DeviceGuard dg(g->dev[j]);
                FR_CUDA(cudaStreamWaitEvent(g->stream[j], g->ev_use[j], 0));
                if (has_root && j == 0) FR_CUDA(cudaMemcpyAsync(g->q[0].p, pin_q, qb, cudaMemcpyHostToDevice, g->stream[0]));
            }
            FR_NCCL(g_nccl.GroupStart());
            for (int j = 0; j < g->nl; ++j) {
                int r = g_nccl.Broadcast(g->q[j].p, g->q[j].p, static_cast<size_t>(B) * g->dim, NCCL_FLOAT32, 0, g->comm[j], g->stream[j]);
                if (r != 0) {
                    g_nccl.GroupEnd();
                    return fail(FR_ECUDA, "ncclBroadcast failed: %s", g_nccl.GetErrorString(r));
                }
            }
            FR_NCCL(g_nccl.GroupEnd());
        }
        rc = group_search_enqueue(g, dq.data(), B, k, od.data(), ok.data(), g->stream.data());
        if (rc != FR_OK) return rc;
        DeviceGuard dg(g->dev[0]);
        FR_CUDA(cudaMemcpyAsync(pin_dist, od[0], db, cudaMemcpyDeviceToHost, g->stream[0]));
        FR_CUDA(cudaMemcpyAsync(pin_keys, ok[0], kb, cudaMemcpyDeviceToHost, g->stream[0]));
        FR_CUDA(cudaEventRecord(g->ev_use[0], g->stream[0]));
        FR_CUDA(cudaEventRecord(slot.done, g->stream[0]));
    }
    {
        // stream 0's merge waited for every shard's lists, so every device has also finished READING this slot's queries
        DeviceGuard dg(g->dev[0]);
        std::shared_lock<std::shared_mutex> wait_lk(fr::graph_wait_mutex());  // not beside another object's graph work
        FR_CUDA(cudaEventSynchronize(slot.done));
    }
    std::memcpy(out_dist, pin_dist, db);
    std::memcpy(out_keys, pin_keys, kb);
    return FR_OK;
}

}  // extern "C"
