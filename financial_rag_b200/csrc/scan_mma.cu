// K2 scan_topk_mma -- tensor-core exact scan with a fused top-k epilogue (sm_100a: tcgen05 + TMEM + TMA).
//
// Replaces the arithmetic behind chromadb Collection.query (parent_child/chroma_child_store.py:63,
// parent_child/multivector_store.py:151) for query batches too large for the CUDA-core stream
// kernel (K1 is FFMA-bound above ~2 queries per pass).
//
// The scan of a query block against the corpus IS a dense contraction, so it runs on the 5th-gen
// tensor cores.  Two instances of one kernel template:
//   CG = 1 (<= 128 queries per pass):  D[128 queries x 128 rows] += A[128 x 16] * B[128 x 16]^T
//   CG = 2 (<= 256 queries per pass):  a CTA pair (cluster of 2, tcgen05 cta_group::2) computes
//          D[256 queries x 256 rows]; each CTA keeps ITS 128 queries and loads ITS half of the
//          corpus tile, so one byte of corpus fetched from HBM feeds twice the math -- at 256
//          queries per pass the kernel sits on the tensor roofline instead of the HBM one.
//   * A = the query block (bf16 copy of the normalised queries, zero-padded), loaded once by TMA
//     into shared memory as six K-chunks of [128 rows x 64 elements], SWIZZLE_128B, K-major;
//   * B = corpus tiles, streamed by TMA ([64 x 128] boxes of the row-major bf16 corpus -- its
//     natural layout is already the K-major operand) through an mbarrier ring;
//   * one elected thread (of the leader CTA when CG = 2) issues 24 tcgen05.mma (K = 16 each) per
//     corpus tile into a TMEM accumulator (4 x 128 columns, or 2 x 256); tcgen05.commit frees the
//     smem stage / publishes the accumulator (multicast to both CTAs of a pair);
//   * one TMEM lane = one query: each of the 128 epilogue threads reads ITS query's scores with
//     tcgen05.ld (64 columns at a time) and keeps a running maximum; only when some lane's maximum
//     beats its current threshold does the warp enter the insert path (transpose 16 columns through
//     shared memory, warp-cooperative sorted insert into that query's list in shared memory).
//     No score ever goes to HBM.
//   * thresholds are shared between CTAs.  A CTA alone only knows the k'-th best of ITS rows, which
//     rises 148x slower than the k'-th best of everything scanned so far.  So the CTAs are dealt
//     into k' groups; slot[group][query] (global, atomicMax of the order-preserving score bits)
//     holds the best score any CTA of the group has INSERTED for the query.  The minimum over the
//     k' slots is a score that k' distinct live rows reach, i.e. a lower bound on the final k'-th
//     selection score, and every CTA gates on max(own k'-th, that minimum).  A row is therefore
//     only dropped when its selection score is <= a value that is itself <= the final k'-th
//     selection score -- all the certification below needs.
// Roofline: at <= 128 queries per pass the MMA work per tile (24 x 64 cycles) is far below the HBM
// time of the tile (128 rows x 768 B): HBM-bound like K1, but for 128 queries at once.  At 256
// queries per pass (CG = 2) MMA time per 256-row tile (24 x 128 cycles per pair) matches its HBM time.
//
// Exactness.  The tensor cores need bf16 queries (q16 below); the reference semantics (and K1) use
// fp32 queries.  So K2 only SELECTS: it keeps k' = 32*KPL > k candidates per query, ranked by the
// q16-query score; rescore_kernel then recomputes the k' scores with the fp32 query (fp32 FMA), sorts them
// by the exact key and CERTIFIES the top-k:  every row outside the candidate set has
//     exact score <= (k'-th selection score) + |q - q16|_2 * |c|_2  (+ fp32 accumulation slack),
// so if the k-th exact score is above that bound the result equals the fp32-query scan; otherwise
// the query is flagged and re-scanned by the stream kernel (scan_stream_fallback_kernel).
#include "fr_kernels.h"
#include "mma_common.cuh"

namespace fr {
namespace mma {

constexpr int M_TILE = 128;                    // queries per CTA (TMEM lanes)
constexpr int N_TILE = TILE_ROWS_CTA;          // corpus rows loaded per CTA per tile
constexpr int CHUNK_BYTES = STAGE_BYTES;       // 16 KB: one [128 x 64] bf16 box (A chunk or B stage)
constexpr int TMEM_COLS = 512;                 // the whole tensor memory: 4 x 128 or 2 x 256 columns
constexpr int THREADS = 192;                   // warp 0 TMA, warp 1 MMA + TMEM alloc, warps 2-5 epilogue

// unsorted pending candidates per query between compactions: 16 beside shared-memory lists, 32 (a full warp sort per
// compaction, half as many compactions against L2) where the lists live in global memory and shared memory has room
#ifndef FR_K2_PEND_K32
#define FR_K2_PEND_K32 16   // experiment knobs (scripts/build_variant.py): pending slots and ring depth at k' = 32
#endif
#ifndef FR_K2_STAGES_K32
#define FR_K2_STAGES_K32 5
#endif
__host__ __device__ constexpr int pend_slots(int kpl) { return kpl >= 4 ? 32 : (kpl == 1 ? FR_K2_PEND_K32 : 16); }
constexpr int RETRY_MAX = M_TILE;              // uncertified queries that get a second tensor-core pass
constexpr int RESCORE_MAX_DIM = 1024;          // widest vectors the rescoring kernel stages in shared memory

// Per-query candidate lists hold CAP = 32*KPL = k' entries.  For k' <= 64 they live in shared memory;
// for k' = 128 (k up to 100) 128 queries x 1 KB would not fit beside the 96 KB query block, so the CTA
// works directly in ITS slice of the `partials` array in global memory (L2-resident, touched only on
// the rare compactions, and each lane only ever touches the entries congruent to its lane id).
template <int KPL>
struct SmemPlan {
    static constexpr bool LISTS_IN_SMEM = KPL <= 2;
    static constexpr int STAGES = (KPL == 1) ? FR_K2_STAGES_K32 : (KPL == 2 ? 3 : 6);  // KPL 4 and 8: lists are not in smem
    static constexpr int CAP = 32 * KPL;
    static constexpr size_t A_OFF = 0;
    static constexpr size_t B_OFF = A_OFF + size_t(K_CHUNKS) * CHUNK_BYTES;
    static constexpr size_t LIST_OFF = B_OFF + size_t(STAGES) * CHUNK_BYTES;   // [128 queries][CAP] sorted
    static constexpr size_t PEND_OFF = LIST_OFF + (LISTS_IN_SMEM ? size_t(M_TILE) * CAP * 8 : 0);  // [4 warps][PEND][32] swizzled
    static constexpr int PEND = pend_slots(KPL);
    static constexpr size_t BAR_OFF = PEND_OFF + size_t(M_TILE) * PEND * 8;
    static constexpr size_t TOTAL = BAR_OFF + 256;
    static constexpr size_t ALLOC = TOTAL + 1024;  // slack to align the base to 1024 B
};
static_assert(SmemPlan<1>::ALLOC <= 232448 && SmemPlan<2>::ALLOC <= 232448 && SmemPlan<4>::ALLOC <= 232448 &&
                  SmemPlan<8>::ALLOC <= 232448,
              "exceeds 227 KB of shared memory");

// v[i] for a run-time i: registers cannot be indexed, so a 6-level tree of selects (63 of them)
__device__ __forceinline__ float mux64(const float (&v)[64], int i) {
    float a[32], b[16], c[8], d[4], e[2];
#pragma unroll
    for (int j = 0; j < 32; ++j) a[j] = (i & 1) ? v[2 * j + 1] : v[2 * j];
#pragma unroll
    for (int j = 0; j < 16; ++j) b[j] = (i & 2) ? a[2 * j + 1] : a[2 * j];
#pragma unroll
    for (int j = 0; j < 8; ++j) c[j] = (i & 4) ? b[2 * j + 1] : b[2 * j];
#pragma unroll
    for (int j = 0; j < 4; ++j) d[j] = (i & 8) ? c[2 * j + 1] : c[2 * j];
#pragma unroll
    for (int j = 0; j < 2; ++j) e[j] = (i & 16) ? d[2 * j + 1] : d[2 * j];
    return (i & 32) ? e[1] : e[0];
}

// Fold the pending candidates of one query into its sorted list (all 32 lanes cooperate).
//   list: [32*KPL] sorted descending (0 = empty), pend_w: this warp's pending block, n = pending count.
// Returns the new k'-th key (0 while the list is not full).
template <int KPL>
__device__ __forceinline__ uint64_t compact_query(uint64_t *list, const uint64_t *pend_w, int tq, int n, int lane) {
    uint64_t p = (lane < n) ? pend_w[lane * 32 + ((tq + lane) & 31)] : 0ull;
    p = bitonic_sort32_desc(p, lane);
    uint64_t L[KPL];
#pragma unroll
    for (int j = 0; j < KPL; ++j) L[j] = list[j * 32 + lane];
    fold_sorted32<KPL>(L, p, lane);  // push the sorted pending block down the list, 32 entries at a time
#pragma unroll
    for (int j = 0; j < KPL; ++j) list[j * 32 + lane] = L[j];
    const uint32_t lo = __shfl_sync(FULL_MASK, static_cast<uint32_t>(L[KPL - 1]), 31);
    const uint32_t hi = __shfl_sync(FULL_MASK, static_cast<uint32_t>(L[KPL - 1] >> 32), 31);
    return (static_cast<uint64_t>(hi) << 32) | lo;
}

// W queries at once: the W sorting networks are independent dependency chains in one basic block, so their shuffles
// overlap.  One network is ~20 dependent stages of two shuffles each and issues an instruction every fifth cycle or so;
// in the first tiles of a launch every lane of the warp fills its pending block at the same column, i.e. 32
// compactions are due at once -- taken one by one they were most of the fixed cost of a launch.
// kth[w] = the new k'-th key of query tq[w] (0 while the list is not full).
template <int KPL, int W>
__device__ __forceinline__ void compact_queries(uint64_t *lists, int cap, const uint64_t *pend_w, const int (&tq)[W],
                                                const int (&n)[W], int lane, uint64_t (&kth)[W]) {
    uint64_t p[W];
    uint64_t L[W][KPL];
#pragma unroll
    for (int w = 0; w < W; ++w) {
        p[w] = (lane < n[w]) ? pend_w[lane * 32 + ((tq[w] + lane) & 31)] : 0ull;
#pragma unroll
        for (int j = 0; j < KPL; ++j) L[w][j] = lists[static_cast<size_t>(tq[w]) * cap + j * 32 + lane];
    }
#pragma unroll
    for (int k = 2; k <= 32; k <<= 1) {
#pragma unroll
        for (int j = k >> 1; j > 0; j >>= 1) {
            const bool keep_max = ((lane & j) == 0) == ((lane & k) == 0);
#pragma unroll
            for (int w = 0; w < W; ++w) {
                const uint64_t y = shfl_xor_u64(p[w], j);
                p[w] = keep_max ? umax64(p[w], y) : umin64(p[w], y);
            }
        }
    }
    // the fold of fold_sorted32 (fr_common.cuh), W lists at a time
#pragma unroll
    for (int w = 0; w < W; ++w) L[w][KPL - 1] = umax64(L[w][KPL - 1], reverse32(p[w], lane));
#pragma unroll
    for (int s = KPL / 2; s > 0; s >>= 1) {
#pragma unroll
        for (int w = 0; w < W; ++w) {
#pragma unroll
            for (int j = 0; j < KPL; ++j) {
                if ((j & s) == 0) {
                    const uint64_t hi = umax64(L[w][j], L[w][j + s]), lo = umin64(L[w][j], L[w][j + s]);
                    L[w][j] = hi;
                    L[w][j + s] = lo;
                }
            }
        }
    }
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) {
        const bool up = (lane & s) == 0;
#pragma unroll
        for (int w = 0; w < W; ++w) {
#pragma unroll
            for (int j = 0; j < KPL; ++j) {
                const uint64_t y = shfl_xor_u64(L[w][j], s);
                L[w][j] = up ? umax64(L[w][j], y) : umin64(L[w][j], y);
            }
        }
    }
#pragma unroll
    for (int w = 0; w < W; ++w) {
        if (n[w] > 0) {  // warp-uniform; a way without pending entries (padding of a partial group) leaves its list alone
#pragma unroll
            for (int j = 0; j < KPL; ++j) lists[static_cast<size_t>(tq[w]) * cap + j * 32 + lane] = L[w][j];
        }
        const uint32_t lo32 = __shfl_sync(FULL_MASK, static_cast<uint32_t>(L[w][KPL - 1]), 31);
        const uint32_t hi32 = __shfl_sync(FULL_MASK, static_cast<uint32_t>(L[w][KPL - 1] >> 32), 31);
        kth[w] = (static_cast<uint64_t>(hi32) << 32) | lo32;
    }
}

// how many compactions the epilogue takes at once: bounded by the registers the lists occupy (KPL keys per lane each)
__host__ __device__ constexpr int compact_ways(int kpl) { return kpl <= 2 ? 4 : 2; }

// a[0..N) -> descending, in registers (a bitonic network with compile-time indices: N/4 log2 N (log2 N + 1) FMNMX pairs)
template <int N>
__device__ __forceinline__ void reg_sort_desc(float (&a)[N]) {
#pragma unroll
    for (int k = 2; k <= N; k <<= 1) {
#pragma unroll
        for (int j = k >> 1; j > 0; j >>= 1) {
#pragma unroll
            for (int i = 0; i < N; ++i) {
                const int l = i ^ j;
                if (l > i) {
                    const float hi = fmaxf(a[i], a[l]), lo = fminf(a[i], a[l]);
                    const bool desc = (i & k) == 0;
                    a[i] = desc ? hi : lo;
                    a[l] = desc ? lo : hi;
                }
            }
        }
    }
}

// the 32nd largest of 64 values: sort the halves, then the elementwise maxima of one half against the other reversed are
// the 32 largest of the union, and their minimum is the answer
__device__ __forceinline__ float reg_select32_of64(const float (&g)[64]) {
    float a[32], b[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) {
        a[i] = g[i];
        b[i] = g[32 + i];
    }
    reg_sort_desc<32>(a);
    reg_sort_desc<32>(b);
    float m = fmaxf(a[0], b[31]);
#pragma unroll
    for (int i = 1; i < 32; ++i) m = fminf(m, fmaxf(a[i], b[31 - i]));
    return m;
}

// ---------------------------------------------------------------------------------------------
// grid = P * co * CG CTAs.  P = partial lists per query = corpus "streams": stream p takes corpus
// tiles p, p + P, ... of 128*CG rows.  `co` query groups (of 128*CG queries each) are co-resident:
// CTA unit u = blockIdx.x / CG serves group u % co on stream u / co, so the `co` units that walk the
// same tile sequence sit on neighbouring SMs, run in lock step (same work per tile) and all but the
// first of them find the corpus tile in L2 -- HBM traffic per query drops by `co` in the regime
// where one pass over the corpus is not HBM-bound any more.  partials: [P][nq_total][ksel].
//   nq_launch = queries served by this launch (all groups), q_row0 = first of them.
template <int KPL, int CG>
__global__ void __launch_bounds__(THREADS, 1)
scan_mma_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_c,
                const int64_t *__restrict__ keys_or_null, int64_t n_rows, int nq_launch, int q_row0_launch, int ksel,
                uint64_t *__restrict__ partials, int nq_total, int co, uint32_t *__restrict__ tau_g, int dbg,
                const int *__restrict__ nq_dev, const float *__restrict__ tau0, uint32_t *__restrict__ progress,
                int max_lead, uint32_t *__restrict__ hist_g) {
    using Plan = SmemPlan<KPL>;
    constexpr int STAGES = Plan::STAGES;
    constexpr int CAP = Plan::CAP;
    constexpr int ACC_COLS = N_TILE * CG;            // accumulator width = corpus rows per tile
    constexpr int TMEM_BUFS = TMEM_COLS / ACC_COLS;  // 4 or 2
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t *smem_a = smem + Plan::A_OFF;
    uint8_t *smem_b = smem + Plan::B_OFF;
    uint64_t *lists_smem = reinterpret_cast<uint64_t *>(smem + Plan::LIST_OFF);
    uint64_t *pend_all = reinterpret_cast<uint64_t *>(smem + Plan::PEND_OFF);
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + Plan::BAR_OFF);
    // barrier slots: full[STAGES] | empty[STAGES] | tmem_full[4] | tmem_empty[4] | a_full | tmem_ptr
    // (identical offsets in both CTAs of a pair; full / tmem_empty / a_full are used in the leader only)
    const uint32_t bar_full = smem_u32(bars);
    const uint32_t bar_empty = smem_u32(bars + STAGES);
    const uint32_t bar_tfull = smem_u32(bars + 2 * STAGES);
    const uint32_t bar_tempty = smem_u32(bars + 2 * STAGES + 4);
    const uint32_t bar_afull = smem_u32(bars + 2 * STAGES + 8);
    uint32_t *tmem_ptr_smem = reinterpret_cast<uint32_t *>(bars + 2 * STAGES + 9);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const uint32_t rank = (CG == 2) ? cluster_ctarank() : 0u;  // position inside the CTA pair
    const int unit = blockIdx.x / CG;
    const int grp = unit % co;                       // query group of this CTA (pair)
    const int pair = unit / co;                      // corpus stream
    const int npairs = gridDim.x / (CG * co);
    const int q_row0 = q_row0_launch + grp * (M_TILE * CG);  // first query row of the group
    const int q_offset = q_row0;
    // second-chance launches learn their query count on the device (every CTA reads the same value)
    if (nq_dev != nullptr) {
        nq_launch = *nq_dev;
        if (nq_launch <= 0) return;
    }
    const int nq = min(M_TILE * CG, nq_launch - grp * (M_TILE * CG));

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(bar_full + 8 * s, 1);
            mbar_init(bar_empty + 8 * s, 1);
        }
        for (int b = 0; b < TMEM_BUFS; ++b) {
            mbar_init(bar_tfull + 8 * b, 1);
            mbar_init(bar_tempty + 8 * b, 4 * CG);  // one arrival per epilogue warp of the pair
        }
        mbar_init(bar_afull, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        if constexpr (CG == 1) {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr_smem)),
                         "r"(TMEM_COLS)
                         : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        } else {
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr_smem)),
                         "r"(TMEM_COLS)
                         : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
        }
    }
    // this CTA's 128 candidate lists: shared memory, or (k' = 128) its own slice of `partials`
    uint64_t *lists = Plan::LISTS_IN_SMEM
                          ? lists_smem
                          : partials + (static_cast<size_t>(pair) * nq_total + q_offset + static_cast<int>(rank) * M_TILE) * CAP;
    // epilogue warps clear their queries' lists while the allocation happens
    if (warp >= 2) {
        const int n_clear = Plan::LISTS_IN_SMEM ? M_TILE * CAP : max(0, min(M_TILE, nq - static_cast<int>(rank) * M_TILE)) * CAP;
        for (int i = threadIdx.x - 64; i < n_clear; i += 128) lists[i] = 0ull;
    }
    tc_fence_before();
    __syncthreads();
    if constexpr (CG == 2) cluster_sync_all();  // the peer's barriers are initialised before anyone signals them
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;

    constexpr int TILE_ROWS = N_TILE * CG;
    const int64_t num_tiles = (n_rows + TILE_ROWS - 1) / TILE_ROWS;
    const int nq_local = max(0, min(M_TILE, nq - static_cast<int>(rank) * M_TILE));  // this CTA's live queries

    if (warp == 0) {
        // ===================== TMA producer (both CTAs of a pair) =====================
        if (lane == 0) {
            // barriers the MMA issuer waits on live in the leader CTA
            const uint32_t l_afull = (CG == 2) ? map_to_cta(bar_afull, 0) : bar_afull;
            const uint32_t l_full = (CG == 2) ? map_to_cta(bar_full, 0) : bar_full;
            if (rank == 0) mbar_expect_tx(bar_afull, K_CHUNKS * CHUNK_BYTES * CG);
#pragma unroll
            for (int kc = 0; kc < K_CHUNKS; ++kc)
                tma_load_2d<CG>(smem_u32(smem_a + kc * CHUNK_BYTES), &tmap_q, l_afull, kc * K_CHUNK,
                                q_row0 + static_cast<int>(rank) * M_TILE);
            int stage = 0;
            uint32_t phase = 0;
            // Co-resident groups of one stream only share its tiles through L2 while they stay within a few tiles of
            // each other, and nothing keeps them there: their epilogues do different amounts of insert work, the
            // distances random-walk, and once a pair trails by more than L2 retains (~30 us of streaming) it re-reads
            // the corpus from HBM for the rest of the scan (measured with 4 groups over 100M rows: 233 GB of DRAM reads
            // per launch against 76.8 GB of corpus).  So every pair publishes how many tiles it has requested, and a
            // pair that is more than max_lead tiles ahead of the slowest sibling waits for it.  One wait is bounded
            // (~20 ms of polling): a sibling that makes no progress for that long is not resident, and the pair stops
            // waiting for good -- sharing is an optimisation, never a dependency.
            bool throttle = progress != nullptr && co > 1 && rank == 0;
            uint32_t *my_progress = throttle ? progress + static_cast<size_t>(pair) * co + grp : nullptr;
            const uint32_t *siblings = throttle ? progress + static_cast<size_t>(pair) * co : nullptr;
            uint32_t issued = 0;
            uint32_t slowest = 0;  // last known progress of the slowest sibling: polled only when the lead may be used up
            for (int64_t t = pair; t < num_tiles; t += npairs, ++issued) {
                if (throttle && issued > slowest + static_cast<uint32_t>(max_lead)) {
                    int spins = 0;
                    while (true) {
                        slowest = 0xffffffffu;
                        for (int g = 0; g < co; ++g) {
                            uint32_t x;
                            asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(x) : "l"(siblings + g));
                            slowest = min(slowest, x);
                        }
                        if (slowest == 0xffffffffu) {  // every sibling is done
                            throttle = false;
                            break;
                        }
                        if (slowest + static_cast<uint32_t>(max_lead) >= issued) break;
                        if (++spins > 20000) {
                            throttle = false;
                            break;
                        }
                        __nanosleep(64);
                    }
                }
                // diagnostics: dbg & 8 = every stream re-reads the same 64 tiles (L2-resident): no HBM traffic, same L2->SM traffic
                const int64_t tt = (dbg & 8) ? (t & 63) : t;
                const int row0 = static_cast<int>(tt * TILE_ROWS) + static_cast<int>(rank) * N_TILE;
#pragma unroll 1
                for (int kc = 0; kc < K_CHUNKS; ++kc) {
                    mbar_wait(bar_empty + 8 * stage, phase ^ 1);
                    if (dbg & 1) {  // diagnostics: no corpus traffic at all, the MMAs re-read stale smem
                        if (rank == 0) mbar_arrive(bar_full + 8 * stage);
                    } else {
                        if (rank == 0) mbar_expect_tx(bar_full + 8 * stage, CHUNK_BYTES * CG);
                        tma_load_2d<CG>(smem_u32(smem_b + stage * CHUNK_BYTES), &tmap_c, l_full + 8 * stage, kc * K_CHUNK, row0);
                    }
                    if (++stage == STAGES) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
                if (my_progress != nullptr)
                    asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(my_progress), "r"(issued + 1u) : "memory");
            }
            if (my_progress != nullptr)  // done: nobody waits for this pair any more
                asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(my_progress), "r"(0xffffffffu) : "memory");
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (leader CTA only) =====================
        if (lane == 0 && rank == 0) {
            constexpr uint32_t idesc = make_idesc_bf16(M_TILE * CG, ACC_COLS);
            mbar_wait(bar_afull, 0);
            tc_fence_after();
            int stage = 0;
            uint32_t phase = 0;
            uint32_t it = 0;
            for (int64_t t = pair; t < num_tiles; t += npairs, ++it) {
                const uint32_t buf = it % TMEM_BUFS;
                const uint32_t bphase = (it / TMEM_BUFS) & 1;
                mbar_wait(bar_tempty + 8 * buf, bphase ^ 1);  // both epilogues have drained this accumulator
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + buf * ACC_COLS;
                // Two K-chunks per barrier round trip: the issuing thread waits for two stages, then issues their
                // eight MMAs back to back.  One chunk per wait left ~0.3 us of wait + fence latency exposed between
                // every four MMAs (+2 ... +4 % at every batch size; three chunks per wait starve the 5-stage ring).
                static_assert(K_CHUNKS % 2 == 0, "the MMA loop takes K-chunks in pairs");
#pragma unroll 1
                for (int kc = 0; kc < K_CHUNKS; kc += 2) {
                    const int s0 = stage;
                    const uint32_t p0 = phase;
                    int s1 = stage + 1;
                    uint32_t p1 = phase;
                    if (s1 == STAGES) {
                        s1 = 0;
                        p1 ^= 1;
                    }
                    mbar_wait(bar_full + 8 * s0, p0);
                    if (!(dbg & 32)) mbar_wait(bar_full + 8 * s1, p1);  // diagnostics: dbg & 32 = one chunk per wait
                    tc_fence_after();
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const int sh = h == 0 ? s0 : s1;
                        if (h == 1 && (dbg & 32)) {
                            mbar_wait(bar_full + 8 * s1, p1);
                            tc_fence_after();
                        }
                        const uint32_t a_addr = smem_u32(smem_a + (kc + h) * CHUNK_BYTES);
                        const uint32_t b_addr = smem_u32(smem_b + sh * CHUNK_BYTES);
#pragma unroll
                        for (int k4 = 0; k4 < K_CHUNK / UMMA_K; ++k4) {
                            const uint64_t adesc = make_kmajor_sw128_desc(a_addr + k4 * UMMA_K * 2);
                            const uint64_t bdesc = make_kmajor_sw128_desc(b_addr + k4 * UMMA_K * 2);
                            if (!(dbg & 4))  // diagnostics: dbg & 4 = no MMAs at all
                                tc_mma_bf16<CG>(d_tmem, adesc, bdesc, idesc, ((kc + h) | k4) != 0 ? 1u : 0u);
                        }
                        tc_commit<CG>(bar_empty + 8 * sh);  // smem stage reusable once these MMAs have read it
                    }
                    stage = s1 + 1;
                    phase = p1;
                    if (stage == STAGES) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
                tc_commit<CG>(bar_tfull + 8 * buf);  // accumulator complete
            }
        }
    } else {
        // ===================== epilogue: one TMEM lane = one query =====================
        const int quarter = warp & 3;              // TMEM lane quarter this warp may access
        const int my_q = quarter * 32 + lane;      // query index inside this CTA's block of 128
        uint64_t *my_lists = lists + static_cast<size_t>(quarter) * 32 * CAP;
        constexpr int PEND = Plan::PEND;
        uint64_t *pend_w = pend_all + static_cast<size_t>(warp - 2) * PEND * 32;  // entry j of lane q: [j][(q + j) & 31]
        const bool live = my_q < nq_local;
        // padded query rows never pass the gate; a second-chance query starts at its fixed threshold
        float tau = live ? (tau0 != nullptr ? tau0[q_offset + static_cast<int>(rank) * M_TILE + my_q] : -INFINITY) : INFINITY;
        int cnt = 0;                               // this query's pending (unsorted) candidates
        // shared thresholds: slot[j][query], j < k'.  With at least k' streams, stream s raises slot s % k'
        // to the best score it holds.  With fewer streams, stream s owns the slots s, s + S, s + 2S, ... and
        // raises slot s + r*S to the score of the r-th best row it holds (read off its sorted list after a
        // compaction; r = 0 also eagerly).  Either way every slot is backed by a row of its own, so the
        // minimum over the k' slots is a score that k' distinct live rows reach.
        const int q_cta0 = q_offset + static_cast<int>(rank) * M_TILE;  // first query of this CTA
        const uint32_t *slots_q = tau_g + q_cta0 + (live ? my_q : 0);
        uint32_t *my_slot = tau_g + static_cast<size_t>(pair % ksel) * nq_total + q_cta0 + (live ? my_q : 0);
        // slots this stream owns (tiny corpora with fewer than k'/32 streams leave some slots empty: no sharing)
        const int my_ranks = npairs >= ksel ? 1 : min(32, (ksel - pair + npairs - 1) / npairs);
        // cosine collections: the score histogram of this lane's query, shared by all CTAs (see the refresh below)
        uint32_t *hist_q = (hist_g != nullptr && live && !(dbg & 2048))
                               ? hist_g + static_cast<size_t>(q_cta0 + my_q) * SCORE_HIST_WORDS : nullptr;
        bool hist_on = false;                      // the histogram has produced a bound for this query
        uint32_t best = 0, published = 0;          // order bits of the best score appended / published
        const uint32_t l_tempty = (CG == 2) ? map_to_cta(bar_tempty, 0) : bar_tempty;
        uint32_t it = 0;
        uint32_t next_refresh = 0;
        const uint32_t refresh_cap = (dbg & 16) ? 8u : 64u;  // diagnostics: dbg & 16 = the old dense schedule
        for (int64_t t = pair; t < num_tiles; t += npairs, ++it) {
            const uint32_t buf = it % TMEM_BUFS;
            const uint32_t bphase = (it / TMEM_BUFS) & 1;
            // min over the k' group slots = a score k' distinct rows reach (monotone hints: relaxed loads)
            uint32_t g = 0;
            // Refresh schedule: every tile at first, then at geometrically growing distances (the k'-th best of
            // everything scanned so far moves like 1/rows: a threshold read at 2/3 of the current progress lets at most
            // 1.5 x the candidates through), at least every `refresh_cap` tiles.  k' relaxed loads per lane and refresh are
            // latency the accumulator hand-back waits for: at k' = 128 the fixed every-8-tiles schedule kept the
            // epilogue in the refresh for half of its time.
            const bool refresh_now = it >= next_refresh;
            if (refresh_now) next_refresh = it + 1u + min(it >> 1, refresh_cap - 1u);
            // (once the histogram bounds every live query of the warp -- k' of its rows counted above 1/16 -- the slots
            //  have nothing to add and their k' loads per lane, 32 round trips at k' = 256, are skipped)
            const bool read_slots = !__all_sync(FULL_MASK, !live || hist_on) || (dbg & 4096);
            if (live && refresh_now && read_slots) {
                g = 0xffffffffu;
#pragma unroll 8
                for (int j = 0; j < ksel; ++j) {
                    uint32_t x;
                    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(x) : "l"(slots_q + static_cast<size_t>(j) * nq_total));
                    g = min(g, x);
                }
            }
            // The same refresh reads the query's score histogram (16 coarse + 256 fine counters over [0, 1), filled by
            // every CTA with the rows it appends): the lower edge of the highest bin with k' counted rows at or above it
            // is a score k' distinct live rows reach -- as valid a bound as the slots' minimum, and tighter: the slots give
            // a minimum over per-stream bests, the histogram the k'-th best of everything counted to within a bin.
            // Counters only grow, so a stale or half-updated view only loosens the bound.  Two dependent round trips
            // of four 16-byte loads; the fine counters that lag their coarse bin leave the coarse edge.
            float edge = -INFINITY;
            if (hist_q != nullptr && refresh_now) {
                uint32_t cc[16];
#pragma unroll
                for (int v4 = 0; v4 < 4; ++v4)
                    asm volatile("ld.relaxed.gpu.global.v4.u32 {%0, %1, %2, %3}, [%4];"
                                 : "=r"(cc[4 * v4]), "=r"(cc[4 * v4 + 1]), "=r"(cc[4 * v4 + 2]), "=r"(cc[4 * v4 + 3])
                                 : "l"(hist_q + 4 * v4));
                int cb = -1;
                uint32_t acc = 0u, above = 0u;
#pragma unroll
                for (int b = 15; b >= 1; --b) {  // (bin 0 is never counted)
                    const uint32_t nacc = acc + cc[b];
                    if (cb < 0 && nacc >= static_cast<uint32_t>(ksel)) {
                        cb = b;
                        above = acc;
                    }
                    acc = nacc;
                }
                if (cb >= 1) {
                    uint32_t cf[16];
#pragma unroll
                    for (int v4 = 0; v4 < 4; ++v4)
                        asm volatile("ld.relaxed.gpu.global.v4.u32 {%0, %1, %2, %3}, [%4];"
                                     : "=r"(cf[4 * v4]), "=r"(cf[4 * v4 + 1]), "=r"(cf[4 * v4 + 2]), "=r"(cf[4 * v4 + 3])
                                     : "l"(hist_q + 16 + 16 * cb + 4 * v4));
                    int fb = 0;
                    bool found = false;
#pragma unroll
                    for (int j = 15; j >= 1; --j) {
                        above += cf[j];
                        if (!found && above >= static_cast<uint32_t>(ksel)) {
                            fb = j;
                            found = true;
                        }
                    }
                    edge = static_cast<float>(16 * cb + fb) * (1.0f / 256.0f);
                    hist_on = true;
                }
            }
            mbar_wait(bar_tfull + 8 * buf, bphase);
            tc_fence_after();
            if (g != 0u) tau = fmaxf(tau, unorder_bits(g));
            tau = fmaxf(tau, edge);
            const uint32_t row0 = static_cast<uint32_t>(t * TILE_ROWS);
            const uint32_t tcol = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + buf * ACC_COLS;
            if constexpr (KPL == 1) {
                // ---- first tile: a threshold before anything is inserted.  The lists are empty, so every score of the
                //      tile would pass the gate and go through the pending blocks: ACC_COLS / 16 rounds of 32 compactions
                //      per warp, most of what a launch costs beyond its rows (profiles/r02_sweep_rows_*).  Instead the lane
                //      first reads its query's tile once for the maxima of 64 column groups (of 4 columns, 2 for CG = 1):
                //      the 32nd largest group maximum is reached by 32 DISTINCT rows, so it is a lower bound on the 32nd
                //      best score of the tile, and only the ~40 scores at or above it (instead of all 256 / 128) are walked.
                //      Every row has to be live for that (no tombstones, no zero-filled rows past the end), and a
                //      second-chance pass keeps its fixed threshold. ----
                if (it == 0 && tau0 == nullptr && keys_or_null == nullptr && !(dbg & (2 | 256)) &&  // dbg & 256: A/B switch, answers unchanged
                    (t + 1) * TILE_ROWS <= n_rows) {
                    constexpr int GROUP = ACC_COLS / 64;
                    float gmax[64];
#pragma unroll
                    for (int c = 0; c < ACC_COLS / 64; ++c) {
                        float v[64];
                        tmem_ld64(tcol + c * 64, v);
#pragma unroll
                        for (int gi = 0; gi < 64 / GROUP; ++gi) {
                            float m = v[gi * GROUP];
#pragma unroll
                            for (int e = 1; e < GROUP; ++e) m = fmaxf(m, v[gi * GROUP + e]);
                            gmax[c * (64 / GROUP) + gi] = m;
                        }
                    }
                    // the gate is "score > tau": step one ulp below so that the 32 rows AT the bound pass as well
                    // (+0.0 steps down from -0.0: one step below +0.0 in key order IS -0.0, which compares equal)
                    float bound = reg_select32_of64(gmax);
                    bound = bound == 0.0f ? -0.0f : bound;
                    tau = fmaxf(tau, unorder_bits(order_bits(bound) - 1u));
                }
            }
#pragma unroll 1
            for (int c = 0; c < ACC_COLS / 64; ++c) {
                if (dbg & 2) break;  // diagnostics: accumulators are never read
                float v[64];
                tmem_ld64(tcol + c * 64, v);
                float mx = v[0];
#pragma unroll
                for (int i = 1; i < 64; ++i) mx = fmaxf(mx, v[i]);
                if (!__any_sync(FULL_MASK, mx > tau)) continue;  // the common case
                // ---- candidate path.  Every lane appends its own query's passing scores to its pending
                //      block; a lane whose block fills up gets it folded into its sorted list (warp-wide
                //      bitonic network) and resumes where it stopped, with the raised threshold.
                //      Each lane walks its own pass mask (see below). ----
                const uint32_t rbase = row0 + c * 64;
                int resume = 0;
                while (true) {
                    uint32_t mlo = 0, mhi = 0;
#pragma unroll
                    for (int i = 0; i < 32; ++i) {
                        mlo |= (v[i] > tau) ? (1u << i) : 0u;
                        mhi |= (v[32 + i] > tau) ? (1u << i) : 0u;
                    }
                    uint64_t m = (static_cast<uint64_t>(mhi) << 32) | mlo;
                    m = resume < 64 ? (m & (~0ull << resume)) : 0ull;
                    bool full = false;
                    int stop = 64;
                    // every lane walks ITS OWN passing columns; the score comes out of the lane's registers through a
                    // 6-level select tree (mux64).  The first version walked the UNION of the lanes' columns warp-uniformly
                    // and re-read each one from TMEM: ~4.6 iterations and TMEM round trips per chunk where one lane in
                    // 32 has a candidate -- the dominant fixed cost of a launch (k' ln(rows) candidates per query whatever
                    // the shard size: 0.76 ms per 2048-query launch, profiles/r02_ncu_stall_buckets_scan_mma_short_390k.txt).
                    while (__any_sync(FULL_MASK, m != 0ull && !full)) {
                        if (m != 0ull && !full) {
                            const int i = __ffsll(static_cast<long long>(m)) - 1;
                            const float x = mux64(v, i);
                            const uint32_t row = rbase + i;
                            bool ok = row < n_rows;
                            if (ok && keys_or_null != nullptr) ok = keys_or_null[row] != KEY_TOMBSTONE;
                            if (ok && cnt == PEND) {
                                full = true;  // resume from this column after the compaction
                                stop = i;
                            } else {
                                if (ok) {
                                    pend_w[cnt * 32 + ((lane + cnt) & 31)] = pack_key(x, row);
                                    ++cnt;
                                    best = max(best, order_bits(x));
                                    if (hist_q != nullptr) hist_count(hist_q, x);
                                }
                                m &= m - 1;
                            }
                        }
                    }
                    resume = stop;
                    unsigned fm = __ballot_sync(FULL_MASK, full);
                    if (fm == 0u) break;
                    __syncwarp();
                    // Up to WAYS full blocks are folded at once (see compact_queries).  (Tried and dropped: ways that are
                    // left over taking queries whose block is at least half full -- more folds of less, 390k-row launch
                    // 0.885 -> 0.939 ms, 12.5M rows +1.2 %; profiles/r02_sweep_rows_fixed_cost_ab.jsonl.)
                    while (fm) {
                        constexpr int WAYS = compact_ways(KPL);
                        int tqs[WAYS], ns[WAYS];
                        uint64_t kths[WAYS];
                        int taken = 0;
#pragma unroll
                        for (int w = 0; w < WAYS; ++w) {
                            const bool on = fm != 0u && !(w > 0 && (dbg & 512));  // dbg & 512: one by one (A/B switch)
                            tqs[w] = on ? __ffs(fm) - 1 : (w > 0 ? tqs[0] : 0);
                            ns[w] = on ? PEND : 0;
                            taken += on ? 1 : 0;
                            if (on) fm &= fm - 1;
                        }
                        if (taken == 1) {
                            kths[0] = compact_query<KPL>(my_lists + static_cast<size_t>(tqs[0]) * CAP, pend_w, tqs[0], PEND, lane);
                        } else if (WAYS > 2 && taken <= 2) {
                            const int tq2[2] = {tqs[0], tqs[1]}, n2[2] = {ns[0], ns[1]};
                            uint64_t k2[2];
                            compact_queries<KPL, 2>(my_lists, CAP, pend_w, tq2, n2, lane, k2);
                            kths[0] = k2[0];
                            kths[1] = k2[1];
                        } else {
                            // (a way without a query names query tqs[0] with no pending entries: it stores nothing)
                            compact_queries<KPL, WAYS>(my_lists, CAP, pend_w, tqs, ns, lane, kths);
                        }
#pragma unroll
                        for (int w = 0; w < WAYS; ++w) {
                            if (w >= taken) break;
                            const int tq = tqs[w];
                            if (lane == tq) {
                                cnt = 0;
                                tau = fmaxf(tau, key_threshold(kths[w]));
                            }
                        }
                        if (my_ranks > 1) {  // lane r: entry r of the sorted list (its own store)
#pragma unroll
                            for (int w = 0; w < WAYS; ++w) {
                                if (w >= taken) break;
                                const int tq = tqs[w];
                                if (lane < my_ranks) {
                                    const uint64_t e = my_lists[static_cast<size_t>(tq) * CAP + lane];
                                    if (e != 0ull)
                                        atomicMax(tau_g + static_cast<size_t>(pair + lane * npairs) * nq_total + q_cta0 + quarter * 32 + tq,
                                                  static_cast<uint32_t>(e >> 32));
                                }
                            }
                        }
                    }
                    __syncwarp();
                }
                if (best > published) {  // best live row this CTA holds for the query: other CTAs gate on it
                    published = best;
                    atomicMax(my_slot, best);
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
                if constexpr (CG == 2) mbar_arrive_cluster(l_tempty + 8 * buf);
                else mbar_arrive(bar_tempty + 8 * buf);
            }
        }
        // ---- fold what is still pending, then this CTA's list for each of its queries ----
        __syncwarp();
        {
            constexpr int WAYS = compact_ways(KPL);
            const int q_end = min(32, nq_local - quarter * 32);  // live queries of this warp (<= 0: none)
            for (int q0 = 0; q0 < q_end; q0 += WAYS) {
                int tqs[WAYS], ns[WAYS];
                uint64_t kths[WAYS];
#pragma unroll
                for (int w = 0; w < WAYS; ++w) {
                    const bool on = q0 + w < q_end;
                    tqs[w] = on ? q0 + w : q0;
                    const int n = __shfl_sync(FULL_MASK, cnt, tqs[w]);
                    ns[w] = on ? n : 0;
                }
                compact_queries<KPL, WAYS>(my_lists, CAP, pend_w, tqs, ns, lane, kths);
                __syncwarp();
                if constexpr (Plan::LISTS_IN_SMEM) {
#pragma unroll
                    for (int w = 0; w < WAYS; ++w) {
                        if (q0 + w >= q_end) break;
                        const uint64_t *lp = my_lists + static_cast<size_t>(q0 + w) * CAP;
                        uint64_t *dst = partials + (static_cast<size_t>(pair) * nq_total + q_cta0 + quarter * 32 + q0 + w) * ksel;
                        for (int i = lane; i < ksel; i += 32) dst[i] = lp[i];
                    }
                }
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if constexpr (CG == 2) cluster_sync_all();  // no CTA leaves while its pair may still signal or read it
    if (warp == 1) {
        tc_fence_after();
        if constexpr (CG == 1)
            asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
        else
            asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
    }
}

// ---------------------------------------------------------------------------------------------
// Query preparation for K2 / K2s, one launch: cosine normalisation of the raw queries (bit for bit the
// arithmetic of K0's ingest_kernel, so K1 and K2 see the same prepared query), the bf16 copy the tensor cores
// read (zero-padded to nq_pad rows), the per-query selection error bound |q - bf16(q)|_2, and the reset of the
// call's scratch (threshold slots, failure counters) that would otherwise be three memsets.
// One warp per query row.
__global__ void prep_queries_kernel(const float *__restrict__ raw, int nq, int nq_pad, int dim, float *__restrict__ q_prep,
                                    __nv_bfloat16 *__restrict__ qb, float *__restrict__ err_bound,
                                    float *__restrict__ err_bound_split, float *__restrict__ err_alpha,
                                    float *__restrict__ err_alpha_split, int split, uint32_t *__restrict__ tau_g, int ksel,
                                    uint32_t *__restrict__ hist, int *__restrict__ counters, int n_counters, float bound_scale, int normalize,
                                    const float *__restrict__ cmax, float *__restrict__ inv_scale,
                                    float *__restrict__ qnorm2) {
    const int row = blockIdx.x;
    const int lane = threadIdx.x;  // 32 threads
    if (row == 0 && lane < n_counters) counters[lane] = 0;
    if (row < nq)
        for (int j = lane; j < ksel; j += 32) tau_g[static_cast<size_t>(row) * ksel + j] = 0u;
    if (row < nq && hist != nullptr)
        for (int j = lane; j < SCORE_HIST_WORDS; j += 32) hist[static_cast<size_t>(row) * SCORE_HIST_WORDS + j] = 0u;
    const int c4 = dim >> 2;
    // split: the row holds bf16(q) in columns [0, dim) and bf16(q - bf16(q)) in [dim, 2 dim)
    uint2 *dst = reinterpret_cast<uint2 *>(qb + static_cast<size_t>(row) * dim * (1 + split));
    uint2 *dst_lo = dst + c4;
    if (row >= nq) {
        for (int c = lane; c < c4 * (1 + split); c += 32) dst[c] = make_uint2(0u, 0u);
        return;
    }
    const float4 *s4 = reinterpret_cast<const float4 *>(raw + static_cast<size_t>(row) * dim);
    float ss = 0.0f;
    for (int c = lane; c < c4; c += 32) {
        const float4 v = s4[c];
        ss = fmaf(v.x, v.x, ss);
        ss = fmaf(v.y, v.y, ss);
        ss = fmaf(v.z, v.z, ss);
        ss = fmaf(v.w, v.w, ss);
    }
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) ss += __shfl_xor_sync(FULL_MASK, ss, s);
    // cosine: the query is normalised (K0's arithmetic).  Inner-product collections (FR_IP) score the raw query against
    // raw rows: nothing is normalised, and every absolute error bound below scales with |q| * (largest row norm).
    const float inv = normalize ? 1.0f / (sqrtf(ss) + 1e-30f) : 1.0f;
    float4 *p4 = reinterpret_cast<float4 *>(q_prep + static_cast<size_t>(row) * dim);
    float es = 0.0f, es2 = 0.0f, al = 0.0f, al2 = 0.0f;  // |e|^2 and e . q of the one-term / two-term residual e
    for (int c = lane; c < c4; c += 32) {
        float4 v = s4[c];
        v.x *= inv; v.y *= inv; v.z *= inv; v.w *= inv;
        p4[c] = v;
        const __nv_bfloat162 lo = __floats2bfloat162_rn(v.x, v.y);
        const __nv_bfloat162 hi = __floats2bfloat162_rn(v.z, v.w);
        uint2 o;
        o.x = *reinterpret_cast<const uint32_t *>(&lo);
        o.y = *reinterpret_cast<const uint32_t *>(&hi);
        dst[c] = o;
        const float dx = v.x - __low2float(lo), dy = v.y - __high2float(lo);
        const float dz = v.z - __low2float(hi), dw = v.w - __high2float(hi);
        es = fmaf(dx, dx, es);
        es = fmaf(dy, dy, es);
        es = fmaf(dz, dz, es);
        es = fmaf(dw, dw, es);
        al = fmaf(dx, v.x, al);
        al = fmaf(dy, v.y, al);
        al = fmaf(dz, v.z, al);
        al = fmaf(dw, v.w, al);
        if (split) {  // second bf16 term of the query: what the first one lost (exact differences above)
            const __nv_bfloat162 lo2 = __floats2bfloat162_rn(dx, dy);
            const __nv_bfloat162 hi2 = __floats2bfloat162_rn(dz, dw);
            o.x = *reinterpret_cast<const uint32_t *>(&lo2);
            o.y = *reinterpret_cast<const uint32_t *>(&hi2);
            dst_lo[c] = o;
            const float ex = dx - __low2float(lo2), ey = dy - __high2float(lo2);
            const float ez = dz - __low2float(hi2), ew = dw - __high2float(hi2);
            es2 = fmaf(ex, ex, es2);
            es2 = fmaf(ey, ey, es2);
            es2 = fmaf(ez, ez, es2);
            es2 = fmaf(ew, ew, es2);
            al2 = fmaf(ex, v.x, al2);
            al2 = fmaf(ey, v.y, al2);
            al2 = fmaf(ez, v.z, al2);
            al2 = fmaf(ew, v.w, al2);
        }
    }
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) {
        es += __shfl_xor_sync(FULL_MASK, es, s);
        es2 += __shfl_xor_sync(FULL_MASK, es2, s);
        al += __shfl_xor_sync(FULL_MASK, al, s);
        al2 += __shfl_xor_sync(FULL_MASK, al2, s);
    }
    if (lane == 0) {
        // bound_scale > 1 (option "mma_bound_scale_pct", tests) only makes the certification stricter
        // selection_error_bound() is written for |q| = 1 and rows of norm <= L.  In general, with C the largest row norm,
        // e . c = (e . q^)(c . q^) + e_perp . c_perp, q^ = q / |q|: evaluated at the NORMALISED score t / (|q| C) with
        // E = |e| C and A = |e . q| C / |q| the same formula bounds the selection error in absolute score units.
        const float qn = normalize ? 1.0f : sqrtf(ss);
        const float cm = cmax != nullptr ? *cmax : 1.0f;
        const float scale = fmaxf(qn * cm, 1e-30f);
        const float a_scale = cm / fmaxf(qn, 1e-30f);
        if (inv_scale) inv_scale[row] = 1.0f / scale;
        if (qnorm2) qnorm2[row] = normalize ? 1.0f : ss;
        err_bound[row] = sqrtf(es) * bound_scale * cm;
        err_alpha[row] = fabsf(al) * bound_scale * a_scale;
        if (split) {
            err_bound_split[row] = sqrtf(es2) * bound_scale * cm;
            err_alpha_split[row] = fabsf(al2) * bound_scale * a_scale;
        }
    }
    (void)nq_pad;
}

// ---------------------------------------------------------------------------------------------
// How far the selection score of a row (the score of the bf16 query terms the tensor cores read) can fall below
// its exact fp32-query score s, for any stored row c (|c| <= L: a bf16-rounded unit vector) with s >= t.
// With e = q - (what the scan read), |q| = 1 and c = s q + c_perp:   e . c = (e . q) s + e_perp . c_perp,
// |c_perp|^2 = |c|^2 - s^2.  So  selection >= s - |e . q| s - |e| sqrt(L^2 - s^2)  as well as the plain
// Cauchy-Schwarz  selection >= s - |e| L;  both right-hand sides grow with s, so every row whose exact score
// reaches t > 0 has a selection score of at least  t - selection_error_bound(t).  For the high scores of real
// embedding neighbourhoods (t ~ 0.6 ... 0.95) the second form is 0.8 ... 0.3 of the first.
__device__ __forceinline__ float selection_error_bound(float t, float e_norm, float e_dot_q) {
    constexpr float L = 1.004f;
    const float plain = e_norm * L;
    if (!(t > 0.0f)) return plain;
    const float perp = sqrtf(fmaxf(L * L - t * t, 0.0f));
    return fminf(plain, e_dot_q * 1.001f * t + e_norm * perp);
}

// ---------------------------------------------------------------------------------------------
// rescore_kernel: exact fp32-query scores of the k' selected rows, exact order, certification.
// One CTA (4 warps) per query.  sel: [B][ksel] packed keys sorted by selection score.
//   first pass : CTA j handles query j; an uncertified query is flagged, appended to fail_list and its k-th
//                exact score kept (kth_exact_out) for the second-chance pass.
//   second pass: CTA j < *limit handles query idx_list[j], whose selection list (sel[j]) was collected above the
//                fixed threshold tau0[j].  A list that did not fill up holds EVERY row above tau0[j]; everything
//                else scores at most tau0[j] + bound, so the query is certified iff its k-th exact score beats that.
template <int KPL, bool F32ROWS>
__global__ void __launch_bounds__(256)
rescore_kernel(const uint64_t *__restrict__ sel, int ksel, const float *__restrict__ queries,
               const uint8_t *__restrict__ corpus, const int64_t *__restrict__ row_keys,
               const float *__restrict__ err_bound, const float *__restrict__ err_alpha, int k, float *__restrict__ out_dist,
               uint64_t *__restrict__ out_packed, int64_t *__restrict__ out_keys, uint8_t *__restrict__ flags,
               int *__restrict__ fail_count, int *__restrict__ fail_list, unsigned long long *__restrict__ fail_total,
               float *__restrict__ kth_exact_out, const int *__restrict__ idx_list, const int *__restrict__ limit,
               const float *__restrict__ tau0, int dim, unsigned long long *__restrict__ fail_total2, float acc_slack,
               float extra_bound, const float *__restrict__ inv_scale, int l2, const float *__restrict__ qnorm2,
               const float *__restrict__ cmax) {
    __shared__ float sq[RESCORE_MAX_DIM];
    __shared__ uint64_t exact[32 * KPL];
    const int j_cta = blockIdx.x;
    if (limit != nullptr && j_cta >= *limit) return;
    const int b = idx_list != nullptr ? idx_list[j_cta] : j_cta;  // the query this CTA answers
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int e = threadIdx.x; e < dim; e += blockDim.x) sq[e] = queries[static_cast<size_t>(b) * dim + e];
    for (int i = threadIdx.x; i < 32 * KPL; i += blockDim.x) exact[i] = 0ull;
    __syncthreads();
    const uint64_t *s = sel + static_cast<size_t>(j_cta) * ksel;
    const int nwarps = blockDim.x >> 5;
    // two candidate rows per step and warp: their loads are issued together (a row is 3 x 256 B per warp at width 384;
    // one row at a time left the warp waiting a full DRAM latency per candidate -- 41 us for 128 candidates)
    for (int j = warp; j < ksel; j += 2 * nwarps) {
        const int j1 = j + nwarps;
        const uint64_t key0 = s[j], key1 = j1 < ksel ? s[j1] : 0ull;  // warp-uniform
        if (key0 == 0ull && key1 == 0ull) continue;
        const uint32_t row0 = key_row(key0 != 0ull ? key0 : key1), row1 = key_row(key1 != 0ull ? key1 : key0);
        float acc0 = 0.0f, acc1 = 0.0f;
        if constexpr (F32ROWS) {  // fp32 collection: the exact score is taken from the fp32 row, 16 bytes per lane and step
            const float4 *rp0 = reinterpret_cast<const float4 *>(corpus + static_cast<size_t>(row0) * (static_cast<size_t>(dim) * 4));
            const float4 *rp1 = reinterpret_cast<const float4 *>(corpus + static_cast<size_t>(row1) * (static_cast<size_t>(dim) * 4));
            for (int c = lane; c < dim / 4; c += 32) {
                const float4 w0 = rp0[c], w1 = rp1[c];
                const float *qq = sq + c * 4;
                acc0 = fmaf(w0.x, qq[0], acc0);
                acc0 = fmaf(w0.y, qq[1], acc0);
                acc0 = fmaf(w0.z, qq[2], acc0);
                acc0 = fmaf(w0.w, qq[3], acc0);
                acc1 = fmaf(w1.x, qq[0], acc1);
                acc1 = fmaf(w1.y, qq[1], acc1);
                acc1 = fmaf(w1.z, qq[2], acc1);
                acc1 = fmaf(w1.w, qq[3], acc1);
            }
        } else {
            // 8-byte pieces (4 bf16) of the row, lane-strided: coalesced 256 bytes per warp step
            const uint2 *rp0 = reinterpret_cast<const uint2 *>(corpus + static_cast<size_t>(row0) * (static_cast<size_t>(dim) * 2));
            const uint2 *rp1 = reinterpret_cast<const uint2 *>(corpus + static_cast<size_t>(row1) * (static_cast<size_t>(dim) * 2));
            for (int c = lane; c < dim / 4; c += 32) {
                const uint2 w0 = rp0[c], w1 = rp1[c];
                const float *qq = sq + c * 4;
                const float a0 = __uint_as_float(w0.x << 16), a1 = __uint_as_float(w0.x & 0xffff0000u);
                const float a2 = __uint_as_float(w0.y << 16), a3 = __uint_as_float(w0.y & 0xffff0000u);
                const float b0 = __uint_as_float(w1.x << 16), b1 = __uint_as_float(w1.x & 0xffff0000u);
                const float b2 = __uint_as_float(w1.y << 16), b3 = __uint_as_float(w1.y & 0xffff0000u);
                if (l2) {  // warp-uniform: squared distance by direct differences (no cancellation), as K1 computes it
                    float d;
                    d = qq[0] - a0; acc0 = fmaf(d, d, acc0);
                    d = qq[1] - a1; acc0 = fmaf(d, d, acc0);
                    d = qq[2] - a2; acc0 = fmaf(d, d, acc0);
                    d = qq[3] - a3; acc0 = fmaf(d, d, acc0);
                    d = qq[0] - b0; acc1 = fmaf(d, d, acc1);
                    d = qq[1] - b1; acc1 = fmaf(d, d, acc1);
                    d = qq[2] - b2; acc1 = fmaf(d, d, acc1);
                    d = qq[3] - b3; acc1 = fmaf(d, d, acc1);
                } else {
                    acc0 = fmaf(a0, qq[0], acc0);
                    acc0 = fmaf(a1, qq[1], acc0);
                    acc0 = fmaf(a2, qq[2], acc0);
                    acc0 = fmaf(a3, qq[3], acc0);
                    acc1 = fmaf(b0, qq[0], acc1);
                    acc1 = fmaf(b1, qq[1], acc1);
                    acc1 = fmaf(b2, qq[2], acc1);
                    acc1 = fmaf(b3, qq[3], acc1);
                }
            }
        }
#pragma unroll
        for (int sft = 16; sft > 0; sft >>= 1) {
            acc0 += __shfl_xor_sync(FULL_MASK, acc0, sft);
            acc1 += __shfl_xor_sync(FULL_MASK, acc1, sft);
        }
        if (lane == 0) {  // l2: the exact key orders by -distance (larger is better, like every other score)
            if (key0 != 0ull) exact[j] = pack_key(l2 ? -acc0 : acc0, row0);
            if (key1 != 0ull) exact[j1] = pack_key(l2 ? -acc1 : acc1, row1);
        }
    }
    __syncthreads();
    if (warp != 0) return;
    WarpTopK<KPL> lst;
    int n_valid = 0;
    if constexpr (KPL == 1) {  // 32 candidates: one key per lane, one bitonic sort
        const uint64_t key = exact[lane];
        n_valid = __popc(__ballot_sync(FULL_MASK, key != 0ull));
        lst.e[0] = bitonic_sort32_desc(key, lane);
    } else {  // sort the candidates block by block and fold the blocks together (bitonic networks, no serial inserts)
        lst.clear();
#pragma unroll
        for (int blk = 0; blk < KPL; ++blk) {
            const uint64_t key = blk * 32 + lane < ksel ? exact[blk * 32 + lane] : 0ull;
            n_valid += __popc(__ballot_sync(FULL_MASK, key != 0ull));
            fold_sorted32<KPL>(lst.e, bitonic_sort32_desc(key, lane), lane);
        }
    }
    // certification: a row outside the candidate set has a selection score <= ceiling (the k'-th selection score of
    // a full list, or the fixed threshold of a second-chance list); had its exact score reached the k-th exact score
    // found here, its selection score would be above (k-th exact) - selection_error_bound.  acc_slack (1e-5 per 384
    // products) covers the fp32 accumulation orders of the two kernels.
    const uint64_t last_sel = s[ksel - 1];
    const float kth_exact = n_valid >= k ? key_score(lst.kth(k)) : -INFINITY;
    // (extra_bound: the scan scored a bf16 copy c16 of an fp32 row c; q . (c - c16) is at most 2^-9 sum|q_i c_i|, so a
    //  row whose exact score reaches t scores at least t - extra_bound against the query on the copy)
    // (inner-product collections: the absolute slacks scale with |q| * max row norm = 1 / inv_scale, and the bound is
    //  evaluated at the normalised score -- see prep_queries_kernel; cosine: inv_scale = 1)
    const float isc = inv_scale != nullptr ? inv_scale[b] : 1.0f;
    const float t_copy = kth_exact - extra_bound / isc;
    float floor_sel = t_copy - selection_error_bound(t_copy * isc, err_bound[b], err_alpha[b]) - acc_slack / isc;
    if (l2) {
        // kth_exact = -(k-th smallest exact distance).  In exact arithmetic 2 q.c - |c|^2 = |q|^2 - d, so every row at least
        // as close has t >= |q|^2 - d_k, and its selection value 2 qb.c - |c|^2 is below t by at most 2 |e . c| <= 2 |e| C
        // (err_bound = |e| C) plus rounding: three fp32 sums of 384 non-negative terms, <= 2.3e-5 (|q| + C)^2 in all.
        const float qn = sqrtf(qnorm2[b]), cm = *cmax;
        floor_sel = (qnorm2[b] + kth_exact) - 2.0f * 1.004f * err_bound[b] - 4e-5f * (qn + cm) * (qn + cm) - acc_slack / isc;
    }
    bool certified = true;
    if (last_sel != 0ull) {                       // list full: ceiling = the k'-th selection score
        certified = n_valid >= k && floor_sel > key_score(last_sel);
    } else if (tau0 != nullptr) {                 // not full, but only rows above tau0 were collected
        certified = n_valid >= k && floor_sel > tau0[j_cta];
    }                                             // not full and no threshold: every live row is a candidate
#pragma unroll
    for (int j = 0; j < KPL; ++j) {
        const int i = j * 32 + lane;
        if (i >= k) continue;
        const uint64_t key = lst.e[j];
        const int64_t o = static_cast<int64_t>(b) * k + i;
        if (key == 0ull) {
            if (out_dist) out_dist[o] = INFINITY;
            if (out_packed) out_packed[o] = 0ull;
            out_keys[o] = -1;
        } else {
            if (out_dist) out_dist[o] = l2 ? -key_score(key) : 1.0f - key_score(key);
            if (out_packed) out_packed[o] = key;
            out_keys[o] = row_keys[key_row(key)];
        }
    }
    if (lane == 0) {
        flags[b] = certified ? 0 : 1;
        if (kth_exact_out) kth_exact_out[b] = kth_exact;
        if (!certified) {
            fail_list[atomicAdd(fail_count, 1)] = b;
            if (fail_total) atomicAdd(fail_total, 1ull);
            if (fail_total2) atomicAdd(fail_total2, 1ull);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Second chance for uncertified queries.  A query fails when its k-th exact score is within the rounding
// bound of its k'-th selection score -- crowded score neighbourhoods (near-duplicate chunks, k close to k').
// Everything that could still belong to its top-k scores above  T = (k-th exact score) - bound  in selection
// terms, and only a handful of rows do.  So the failed queries (up to RETRY_MAX of them) are gathered into
// one more query block and scanned again by the same tensor-core kernel with the threshold FIXED at T from
// the first tile on and a longer list (scan_mma_retry_ksel): the list then holds every row above T unless it
// overflows, which is exactly what the second rescore pass certifies.  Only what still fails (or did not fit the block) goes to the stream kernel.
// One CTA per retry slot; slot j belongs to slice j / RETRY_MAX, which is one more launch of the scan kernel
// (its own query block, thresholds and live count retry_n[slice]; an empty slice's launches exit at once).
__global__ void retry_prep_kernel(const float *__restrict__ queries, const float *__restrict__ err_bound,
                                  const float *__restrict__ err_alpha, float extra_bound,
                                  const float *__restrict__ kth_exact, const int *__restrict__ fail_count,
                                  const int *__restrict__ fail_list, __nv_bfloat16 *__restrict__ qb_retry,
                                  float *__restrict__ tau0, int *__restrict__ retry_n, uint32_t *__restrict__ tau_g_retry,
                                  int ksel, uint8_t *__restrict__ flags, int slices, int *__restrict__ fail_count2,
                                  int *__restrict__ fail_list2, unsigned long long *__restrict__ rescanned_total,
                                  int *__restrict__ host_mirror, const float *__restrict__ inv_scale) {
    const int j = blockIdx.x;  // retry slot
    const int slice = j / RETRY_MAX, jj = j % RETRY_MAX;
    const int lane = threadIdx.x;
    const int nf = *fail_count;
    if (j == 0) {
        // The call enqueued `slices` second-chance blocks (as many as the previous searches needed: the host reads
        // host_mirror, pinned memory, without ever waiting for it).  Failures beyond them go straight to the stream
        // re-scan -- still exact, only slower -- and the next call brings more blocks.
        if (lane == 0 && host_mirror != nullptr) *host_mirror = nf;
        for (int i = slices * RETRY_MAX + lane; i < nf; i += 32) {
            fail_list2[atomicAdd(fail_count2, 1)] = fail_list[i];
            if (rescanned_total) atomicAdd(rescanned_total, 1ull);
        }
    }
    if (jj == 0 && lane == 0) {
        const int left = nf - slice * RETRY_MAX;
        retry_n[slice] = left < 0 ? 0 : (left < RETRY_MAX ? left : RETRY_MAX);
    }
    if (slice * RETRY_MAX >= nf) return;  // nothing of this slice will be read
    uint32_t *tg = tau_g_retry + static_cast<size_t>(slice) * ksel * RETRY_MAX;  // [ksel][RETRY_MAX] per slice
    for (int g = lane; g < ksel; g += 32) tg[static_cast<size_t>(g) * RETRY_MAX + jj] = 0u;
    const bool on = j < nf;
    const int b = on ? fail_list[j] : 0;
    for (int e = lane; e < DIM; e += 32)
        qb_retry[static_cast<size_t>(j) * DIM + e] = __float2bfloat16_rn(on ? queries[static_cast<size_t>(b) * DIM + e] : 0.0f);
    if (lane == 0) {
        const float isc = (on && inv_scale != nullptr) ? inv_scale[b] : 1.0f;
        const float t_copy = on ? kth_exact[b] - extra_bound / isc : 0.0f;
        tau0[j] = on ? t_copy - selection_error_bound(t_copy * isc, err_bound[b], err_alpha[b]) - 2e-5f / isc : INFINITY;
        if (on) flags[b] = 1;  // stays flagged until the second rescore pass certifies it
    }
}

// Largest row norm of rows [0, n) (inner-product collections: the scale of every selection error bound).  Warp per row.
template <bool BF16>
__global__ void row_norm_max_kernel(const uint8_t *__restrict__ rows, int64_t n, int dim, float *__restrict__ out,
                                    float *__restrict__ norm2_out) {
    const int lane = threadIdx.x & 31;
    float mx = 0.0f;
    for (int64_t r = blockIdx.x * static_cast<int64_t>(blockDim.x >> 5) + (threadIdx.x >> 5); r < n;
         r += static_cast<int64_t>(gridDim.x) * (blockDim.x >> 5)) {
        float ss = 0.0f;
        if (BF16) {
            const uint32_t *p = reinterpret_cast<const uint32_t *>(rows + static_cast<size_t>(r) * dim * 2);
            for (int c = lane; c < dim / 2; c += 32) {
                const uint32_t w = p[c];
                const float a = __uint_as_float(w << 16), b = __uint_as_float(w & 0xffff0000u);
                ss = fmaf(a, a, fmaf(b, b, ss));
            }
        } else {
            const float *p = reinterpret_cast<const float *>(rows + static_cast<size_t>(r) * dim * 4);
            for (int c = lane; c < dim; c += 32) ss = fmaf(p[c], p[c], ss);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(FULL_MASK, ss, o);
        mx = fmaxf(mx, ss);
        if (norm2_out != nullptr && lane == 0) norm2_out[r] = ss;
    }
    // (1 + 2^-7): fp32 summation noise, and for fp32 rows the bf16 copy the scan reads may be 2^-9 longer
    if (lane == 0 && mx > 0.0f) atomicMax(reinterpret_cast<unsigned int *>(out), __float_as_uint(sqrtf(mx) * 1.0078125f));
}

}  // namespace mma

cudaError_t launch_row_norm_max(const void *rows, bool bf16, int64_t n, int dim, float *out, float *norm2_out, cudaStream_t s) {
    if (n <= 0) return cudaSuccess;
    const int grid = static_cast<int>(n / 8 + 1 < 148 * 8 ? n / 8 + 1 : 148 * 8);
    if (bf16) mma::row_norm_max_kernel<true><<<grid, 256, 0, s>>>(static_cast<const uint8_t *>(rows), n, dim, out, norm2_out);
    else mma::row_norm_max_kernel<false><<<grid, 256, 0, s>>>(static_cast<const uint8_t *>(rows), n, dim, out, norm2_out);
    count_launch();
    return cudaGetLastError();
}

// ---- host launchers ---------------------------------------------------------------------------
// candidates kept per query: k' > k leaves the certification a margin (k' - k rows may overtake)
// (k' = 128 for k = 100 left 28 rows of margin: 1.4 % of the queries of a Gaussian corpus and nearly all of a
// clustered one failed the first certification, and every block of failures costs a whole extra corpus pass)
int scan_mma_ksel(int k, int wide) {
    return k <= 16 ? 32 : (k <= 32 ? 64 : (k <= 64 ? 128 : (k <= 100 ? (wide ? 256 : 128) : 0)));
}

// queries served by one corpus pass: a single CTA per SM up to 128, CTA pairs (cta_group::2) above
int scan_mma_group(int nq_total) { return nq_total <= mma::M_TILE ? mma::M_TILE : 2 * mma::M_TILE; }

cudaError_t launch_prep_queries(const PrepArgs &a) {
    if (a.dim % 4 != 0 || a.n_counters > 32) return cudaErrorInvalidValue;
    mma::prep_queries_kernel<<<a.nq_pad, 32, 0, a.stream>>>(a.raw, a.nq, a.nq_pad, a.dim, a.q_prep,
                                                           static_cast<__nv_bfloat16 *>(a.qb), a.err_bound, a.err_bound_split,
                                                           a.err_alpha, a.err_alpha_split, a.split, a.tau_g, a.ksel, a.hist,
                                                           a.counters, a.n_counters, a.bound_scale >= 1.0f ? a.bound_scale : 1.0f,
                                                           a.normalize ? 1 : 0, a.cmax, a.inv_scale, a.qnorm2);
    count_launch();
    return cudaGetLastError();
}

namespace {
template <int KPL, int CG>
cudaError_t launch_one(const MmaScanArgs &a, const CUtensorMap &tq, const CUtensorMap &tc, int nq, int q0, int co,
                       int lists) {
    auto kern = mma::scan_mma_kernel<KPL, CG>;
    if (co > 1 && a.progress != nullptr) {  // progress counters of the co-resident groups, [lists][co]
        cudaError_t em = cudaMemsetAsync(a.progress, 0, static_cast<size_t>(lists) * co * sizeof(uint32_t), a.stream);
        if (em != cudaSuccess) return em;
    }
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         static_cast<int>(mma::SmemPlan<KPL>::ALLOC));
    if (e != cudaSuccess) return e;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(static_cast<unsigned>(lists * co * CG));
    cfg.blockDim = dim3(mma::THREADS);
    cfg.dynamicSmemBytes = mma::SmemPlan<KPL>::ALLOC;
    cfg.stream = a.stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CG;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    e = cudaLaunchKernelEx(&cfg, kern, tq, tc, a.keys_or_null, a.n_rows, nq, q0, a.ksel, a.partials, a.nq_total, co,
                           a.tau_g, a.dbg, a.nq_dev, a.tau0, co > 1 ? a.progress : nullptr, a.max_lead, a.hist);
    count_launch();
    return e != cudaSuccess ? e : cudaGetLastError();
}
}  // namespace

// How a call of nq_total queries is laid over the SMs: full launches of `co` co-resident groups on
// `lists` corpus streams each, and (when the number of groups is not a multiple of co) one tail
// launch whose `co_tail` groups spread over `lists_tail` streams, so that no SM idles.
MmaPlan scan_mma_plan(int sm_count, int64_t n_rows, int nq_total, int co_max) {
    MmaPlan p{};
    p.group = scan_mma_group(nq_total);
    const int cg = p.group / mma::M_TILE;
    const int groups = (nq_total + p.group - 1) / p.group;
    p.co = groups < co_max ? groups : co_max;
    if (p.co < 1) p.co = 1;
    const int64_t tile_rows = static_cast<int64_t>(mma::N_TILE) * cg;
    int64_t tiles = (n_rows + tile_rows - 1) / tile_rows;
    if (tiles < 1) tiles = 1;
    auto streams = [&](int co) {
        int64_t units = sm_count / (cg * co);
        if (units < 1) units = 1;
        return static_cast<int>(tiles < units ? tiles : units);
    };
    p.lists = streams(p.co);
    p.co_tail = groups % p.co;
    p.tail_q0 = p.co_tail ? (groups - p.co_tail) * p.group : nq_total;
    p.lists_tail = p.co_tail ? streams(p.co_tail) : 0;
    p.lists_max = p.lists > p.lists_tail ? p.lists : p.lists_tail;
    return p;
}

cudaError_t launch_scan_mma(const MmaScanArgs &a) {
    CUtensorMap tq, tc;
    if (!mma::make_row_major_map(&tq, a.queries_bf16, a.nq_pad, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16) ||
        !mma::make_row_major_map(&tc, a.corpus, a.n_rows, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16))
        return cudaErrorNotSupported;
    const MmaPlan &p = a.plan;
    const int per_launch = p.group * p.co;  // queries served by one full launch
    for (int q0 = 0; q0 < a.nq_total; q0 += per_launch) {
        const bool tail = q0 >= p.tail_q0;
        const int nq = (a.nq_total - q0 < per_launch) ? (a.nq_total - q0) : per_launch;
        const int co = tail ? p.co_tail : p.co;
        const int lists = tail ? p.lists_tail : p.lists;
        cudaError_t e;
        if (p.group == mma::M_TILE)
            e = a.ksel <= 32    ? launch_one<1, 1>(a, tq, tc, nq, q0, co, lists)
                : a.ksel <= 64  ? launch_one<2, 1>(a, tq, tc, nq, q0, co, lists)
                : a.ksel <= 128 ? launch_one<4, 1>(a, tq, tc, nq, q0, co, lists)
                                : launch_one<8, 1>(a, tq, tc, nq, q0, co, lists);  // second-chance pass of k' = 128
        else
            e = a.ksel <= 32    ? launch_one<1, 2>(a, tq, tc, nq, q0, co, lists)
                : a.ksel <= 64  ? launch_one<2, 2>(a, tq, tc, nq, q0, co, lists)
                : a.ksel <= 128 ? launch_one<4, 2>(a, tq, tc, nq, q0, co, lists)
                                : launch_one<8, 2>(a, tq, tc, nq, q0, co, lists);
        if (e != cudaSuccess) return e;
    }
    return cudaSuccess;
}

cudaError_t launch_rescore(const RescoreArgs &a) {
    if (a.B <= 0) return cudaSuccess;
    if (a.dim > mma::RESCORE_MAX_DIM || a.dim % 4 != 0) return cudaErrorInvalidValue;
#define FR_RESCORE_T(KPL, F32)                                                                                 \
    mma::rescore_kernel<KPL, F32><<<a.B, 256, 0, a.stream>>>(a.sel, a.ksel, a.queries, a.corpus, a.row_keys,        \
                                                            a.err_bound, a.err_alpha, a.k, a.out_dist, a.out_packed, \
                                                            a.out_keys, a.flags, a.fail_count, a.fail_list,        \
                                                            a.fail_total, a.kth_exact, a.idx_list, a.limit, a.tau0, \
                                                            a.dim, a.fail_total2,                                   \
                                                            1e-5f * static_cast<float>(((a.dim + 383) / 384) * (1 + a.split)), \
                                                            a.extra_bound, a.inv_scale, a.l2, a.qnorm2, a.cmax)
#define FR_RESCORE(KPL)                  \
    do {                                 \
        if (a.f32_rows) FR_RESCORE_T(KPL, true); \
        else FR_RESCORE_T(KPL, false);   \
    } while (0)
    if (a.ksel <= 32) FR_RESCORE(1);
    else if (a.ksel <= 64) FR_RESCORE(2);
    else if (a.ksel <= 128) FR_RESCORE(4);
    else FR_RESCORE(8);
#undef FR_RESCORE_T
#undef FR_RESCORE
    count_launch();
    return cudaGetLastError();
}

int scan_mma_retry_max() { return mma::RETRY_MAX; }
// the second-chance pass needs MORE room than the first (the first failed because at least k' rows sit
// above the threshold): four times the candidates for k' = 32, twice for 64 and 128
int scan_mma_retry_ksel(int ksel) { return ksel >= 128 ? 256 : 128; }

cudaError_t launch_retry_prep(const RetryPrepArgs &a) {
    if (a.slices <= 0) return cudaSuccess;
    mma::retry_prep_kernel<<<a.slices * mma::RETRY_MAX, 32, 0, a.stream>>>(a.queries, a.err_bound, a.err_alpha, a.extra_bound, a.kth_exact, a.fail_count,
                                                                         a.fail_list, static_cast<__nv_bfloat16 *>(a.qb_retry),
                                                                         a.tau0, a.retry_n, a.tau_g_retry, a.ksel, a.flags, a.slices, a.fail_count2,
                                                                         a.fail_list2, a.rescanned_total, a.host_mirror, a.inv_scale);
    count_launch();
    return cudaGetLastError();
}

}  // namespace fr
