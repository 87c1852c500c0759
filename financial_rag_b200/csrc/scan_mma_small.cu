// K2s scan_topk_mma_small -- the tensor-core scan for SMALL query batches (1 ... 64 queries), operands swapped.
//
// Replaces the arithmetic behind chromadb Collection.query (parent_child/chroma_child_store.py:63,
// parent_child/multivector_store.py:151) in the HBM-bound regime, where the scan must cost nothing but the
// corpus read.  Measured on B200 (profiles/r01_ablate_small_batch.jsonl): the TMA-fed ring alone streams the
// corpus at 7.2 TB/s with the board already near its 1 kW cap; the [128 queries x 128 rows] MMAs of K2 pull
// the clocks down and cost 17-27 % of that bandwidth -- even when 120 of the 128 query rows are zero padding.
// So here the CORPUS tile is the M operand and the QUERIES are the N operand:
//     D[128 rows x NQ queries] += A[128 rows x 16] * B[NQ queries x 16]^T        (NQ = 16, 32 or 64)
// and the tensor work (and its energy) shrinks with the batch: 1/8 of K2's at 16 queries.
//   * A = corpus tiles, streamed by TMA exactly as in K2 ([64 x 128] boxes, SWIZZLE_128B, K-major) through
//     an mbarrier ring;  B = the query block, loaded once ([64 x NQ] boxes, six K-chunks);
//   * accumulators: 4 x NQ TMEM columns; one TMEM lane = one corpus ROW, one column = one query
//     (SPLIT: two columns, the bf16 hi and lo terms of the query, added in the epilogue -- see the kernel);
//   * epilogue (4 warps, a thread per row): tcgen05.ld its NQ scores, release the accumulator at once, then
//     max_q (score_q - tau_q) > 0 ?  -- NQ FADDs + NQ/2 FMNMX per row; only then the insert path: the queries
//     with a passing row are walked one by one, a ballot over the 32 rows, warp-cooperative sorted insert into
//     that warp's list for the query (shared memory), new threshold.  A warp's first tile is loaded in bulk
//     (one bitonic sort per query) instead of 32 x NQ single inserts.  No score ever goes to HBM;
//   * thresholds are shared between CTAs through tau_g slots as in K2 (slot = stream % k'), laid out
//     [query][slot] so a warp refreshes a query with one coalesced read and a warp-min;
//   * cosine collections: the CTAs also share a score HISTOGRAM per query (hist_g: 16 coarse + 256 fine counters over
//     [0, 1)): every row that passes a gate is counted, and the edge of the highest bin with k' rows at or above it is
//     a score that k' distinct rows reach -- as valid a threshold as the slots' minimum and several times tighter (the
//     slots give the minimum over ~k' CTA maxima, ~5.4/n of a CTA's n rows pass; the true k'-th best of all P n rows
//     lets 0.87/n pass).  With the slots alone 64 queries x top-50 ran 0.28 ms over the 1.10 ms corpus read at 10M
//     rows; started from the final thresholds of an identical search the same launch runs at the read
//     (profiles/r02_k2s_wide_lists_10m.jsonl);
//   * at the end the four warps' lists of each query are merged and written to `partials` in K2's format, so
//     everything downstream (K3 merge, exact rescoring + certification, second chance, stream re-scan) is shared.
// Roofline: HBM.  Algorithmic bytes per launch = rows * 768.
#include "fr_kernels.h"
#include "mma_common.cuh"

namespace fr {
namespace mma {

// warp 0 TMA, warp 1 MMA + TMEM alloc, then EW epilogue warps: 4 (one per TMEM lane quarter, all NQ queries each) for
// 16 queries, 8 (two per quarter, NQ/2 queries each) above -- an epilogue warp is a serial, latency-bound instruction
// stream (tcgen05.ld -> wait -> gate -> rare inserts) whose length grows with the queries it serves; at 32 and 64
// queries four of them could not keep up with the corpus stream.
__host__ __device__ constexpr int small_epilogue_warps(int nq) { return nq <= 16 ? 4 : 8; }
constexpr int S_TMEM_COLS = 256; // 4 accumulators at a 64-column stride
constexpr int S_ACC_STRIDE = 64;
constexpr int S_TMEM_BUFS = 4;

// Shared-memory plan.  The vector width is a run-time property (k_chunks = dim / 64: 6 for the 384-d
// bge-small / gte-small embeddings, 12 for 768-d bert-base token vectors of the multi-vector store), so the
// offsets are computed, identically, on the host (launch size) and in the kernel.
constexpr int S_MAX_STAGES = 8;
// Where the candidate lists live (k' = 32 KPL entries per query):
//   LM_WARP   while they fit in 64 KB (k' = 32; k' = 64 up to 32 queries): one list per (epilogue warp, query) in shared
//             memory, merged at the end -- no sharing, no locks;
//   LM_SHARED beyond (k' = 128 / 256 -- k up to 100: cfg3's per-collection top-50, cfg5's top-100 -- and 64 queries at
//             k' = 64): per-warp lists would be 128 KB+ and leave no room for the ring, so the four lane-quarter warps
//             that serve a query share ONE list per query in shared memory (<= 64 KB) under a per-query spin lock; a
//             warp folds all its passing rows of a tile into it in one bitonic network;
//   (not served) 64 queries x k' = 256 is 128 KB even when shared.  Lists in global memory were tried: the corpus stream
//             evicts them from L2, every insert is a DRAM round trip -- 0.75 of the copy peak at k' = 128
//             (profiles/r02_ncu_full_scan_mma_small_64q_k128_10m_global_lists.txt) and 8 x the roofline time at k' = 256
//             (profiles/r02_launches_k100_b1024_25m.txt).  Batches of 33-64 with k > 64 take K2 instead.
enum { LM_WARP = 0, LM_SHARED = 1, LM_NONE = 2 };
__host__ __device__ constexpr int small_list_mode(int nq, int kpl) {
    return (kpl <= 2 && 4 * nq * 32 * kpl * 8 <= 65536) ? LM_WARP : (nq * 32 * kpl * 8 <= 65536 ? LM_SHARED : LM_NONE);
}
template <int NQ, int KPL, int SPLIT>
struct SmallPlan {
    static constexpr int LMODE = small_list_mode(NQ, KPL);
    static constexpr int CAP = 32 * KPL;
    static constexpr int N_MMA = NQ * (1 + SPLIT);                               // operand rows: hi terms, then lo terms
    static_assert(N_MMA <= 64, "an accumulator is 64 TMEM columns");
    static constexpr size_t Q_CHUNK = size_t(N_MMA) * K_CHUNK * 2;               // [N_MMA x 64] bf16
    static constexpr int EW = small_epilogue_warps(NQ);                          // epilogue warps
    static constexpr int NQH = NQ * 4 / EW;                                      // queries per epilogue warp
    static constexpr int THREADS = 64 + 32 * EW;
    static constexpr size_t LIST_ELEMS = LMODE == LM_SHARED ? size_t(NQ) * CAP          // [NQ][CAP] per CTA
                                                            : size_t(EW) * NQH * CAP;   // [EW warps][NQH][CAP] per CTA
    static constexpr size_t LIST_BYTES = LIST_ELEMS * 8 + size_t(NQ) * 4 * 2;   // + a spin lock and a shared threshold per query
    static constexpr size_t STASH_BYTES = size_t(EW) * NQH * 32 * 4;             // [EW warps][NQH][32 rows] fp32
    static_assert(Q_CHUNK % 1024 == 0, "query chunks must keep the 1024-byte swizzle alignment");
    size_t q_bytes, ring_off, list_off, stash_off, bar_off, alloc;
    int stages;
    __host__ __device__ explicit SmallPlan(int k_chunks) {
        q_bytes = size_t(k_chunks) * Q_CHUNK;                                    // a multiple of 1024
        const size_t fixed = q_bytes + LIST_BYTES + STASH_BYTES + 256 + 1024;    // + barriers + alignment slack
        const long fit = fixed < 232448 ? long((232448 - fixed) / STAGE_BYTES) : 0;
        stages = fit > S_MAX_STAGES ? S_MAX_STAGES : int(fit);
        ring_off = q_bytes;
        list_off = ring_off + size_t(stages) * STAGE_BYTES;
        stash_off = list_off + LIST_BYTES;
        bar_off = stash_off + STASH_BYTES;
        alloc = bar_off + 256 + 1024;
    }
};
constexpr int S_MIN_STAGES = 4;  // a shallower ring cannot cover the HBM latency

// 16 consecutive fp32 columns of this thread's TMEM lane (no wait: the caller waits once for all its loads)
__device__ __forceinline__ void tmem_ld16_nowait(uint32_t taddr, uint32_t *r) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
}

// grid = P CTAs (P = partial lists per query); CTA p takes corpus tiles p, p + P, ... of 128 rows.
// partials: [P][nq_total][ksel].
// SPLIT: the queries are read as TWO bf16 terms, q ~ hi + lo with lo = bf16(q - bf16(q)): the operand holds the NQ
// hi rows followed by the NQ lo rows (one MMA of N = 2 NQ per K step, not two of N = NQ: the cost of an MMA here is
// mostly per instruction), and the epilogue adds the two columns of a query.  Selection error ~1e-5 instead of ~1e-3.
template <int NQ, int KPL, int SPLIT>
__global__ void __launch_bounds__(64 + 32 * small_epilogue_warps(NQ), 1)
scan_mma_small_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_c,
                      const int64_t *__restrict__ keys_or_null, int64_t n_rows, int nq, int ksel,
                      uint64_t *__restrict__ partials, int nq_total, uint32_t *__restrict__ tau_g, int k_chunks,
                      int q0, const int *__restrict__ nq_dev, const float *__restrict__ norm2, int dbg,
                      uint32_t *__restrict__ hist_g) {
    // dbg (option "mma_debug", diagnostics: results are wrong): 2 = the gate runs but no row ever enters a list, 64 = no
    // threshold refresh from the other CTAs
    // norm2 (l2 collections): |c|^2 of every row.  The squared distance |q - c|^2 = |q|^2 - 2 q.c + |c|^2 is smallest where
    // 2 q.c - |c|^2 is largest, so the epilogue turns each dot product into that with ONE FFMA per score and everything
    // downstream (gate, thresholds, lists) works on it unchanged.
    // q0: first query (row of the query block, index into partials / tau_g) this launch serves; with nq_dev the
    // live count comes from the device (retry slices: *nq_dev queries in all, this slice takes [q0, q0 + nq))
    if (nq_dev != nullptr) {
        const int left = *nq_dev - q0;
        nq = left < nq ? left : nq;
        if (nq <= 0) return;  // every thread of every CTA takes the same branch
    }
    using Plan = SmallPlan<NQ, KPL, SPLIT>;
    constexpr int N_MMA = Plan::N_MMA;
    constexpr int EW = Plan::EW;
    constexpr int NQH = Plan::NQH;
    const Plan plan(k_chunks);
    const int STAGES = plan.stages;
    constexpr int CAP = Plan::CAP;
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t *smem_q = smem;
    uint8_t *smem_ring = smem + plan.ring_off;
    constexpr int LMODE = Plan::LMODE;
    static_assert(LMODE != LM_NONE, "this instance is not built");
    uint64_t *lists = reinterpret_cast<uint64_t *>(smem + plan.list_off);
    int *locks = reinterpret_cast<int *>(lists + Plan::LIST_ELEMS);  // LM_SHARED only: one spin lock per query
    float *tau_s = reinterpret_cast<float *>(locks + NQ);            // thresholds as of the last refresh, shared by a group's warps
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + plan.bar_off);
    // barrier slots: full[8] | empty[8] | tmem_full[4] | tmem_empty[4] | q_full | tmem_ptr
    const uint32_t bar_full = smem_u32(bars);
    const uint32_t bar_empty = smem_u32(bars + S_MAX_STAGES);
    const uint32_t bar_tfull = smem_u32(bars + 2 * S_MAX_STAGES);
    const uint32_t bar_tempty = smem_u32(bars + 2 * S_MAX_STAGES + 4);
    const uint32_t bar_qfull = smem_u32(bars + 2 * S_MAX_STAGES + 8);
    uint32_t *tmem_ptr_smem = reinterpret_cast<uint32_t *>(bars + 2 * S_MAX_STAGES + 9);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int cta = blockIdx.x;
    const int ncta = gridDim.x;

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(bar_full + 8 * s, 1);
            mbar_init(bar_empty + 8 * s, 1);
        }
        for (int b = 0; b < S_TMEM_BUFS; ++b) {
            mbar_init(bar_tfull + 8 * b, 1);
            mbar_init(bar_tempty + 8 * b, EW);  // one arrival per epilogue warp
        }
        mbar_init(bar_qfull, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr_smem)),
                     "r"(S_TMEM_COLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (warp >= 2) {
        for (int i = threadIdx.x - 64; i < static_cast<int>(Plan::LIST_ELEMS); i += 32 * EW) lists[i] = 0ull;
        for (int i = threadIdx.x - 64; i < NQ; i += 32 * EW) {
            locks[i] = 0;
            tau_s[i] = -INFINITY;
        }
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;
    const int64_t num_tiles = (n_rows + TILE_ROWS_CTA - 1) / TILE_ROWS_CTA;

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            mbar_expect_tx(bar_qfull, static_cast<uint32_t>(plan.q_bytes));
            for (int kc = 0; kc < k_chunks; ++kc) {  // a query row: columns [0, dim) = bf16(q), [dim, 2 dim) = the lo term
                tma_load_2d<1>(smem_u32(smem_q + kc * Plan::Q_CHUNK), &tmap_q, bar_qfull, kc * K_CHUNK, q0);
                if (SPLIT)  // rows NQ .. 2 NQ - 1 of the operand (NQ * 128 B is a multiple of the 1024-byte swizzle atom)
                    tma_load_2d<1>(smem_u32(smem_q + kc * Plan::Q_CHUNK + NQ * K_CHUNK * 2), &tmap_q, bar_qfull,
                                   (k_chunks + kc) * K_CHUNK, q0);
            }
            int stage = 0;
            uint32_t phase = 0;
            for (int64_t t = cta; t < num_tiles; t += ncta) {
                const int row0 = static_cast<int>(t * TILE_ROWS_CTA);
#pragma unroll 1
                for (int kc = 0; kc < k_chunks; ++kc) {
                    mbar_wait(bar_empty + 8 * stage, phase ^ 1);
                    mbar_expect_tx(bar_full + 8 * stage, STAGE_BYTES);
                    tma_load_2d<1>(smem_u32(smem_ring + stage * STAGE_BYTES), &tmap_c, bar_full + 8 * stage, kc * K_CHUNK, row0);
                    if (++stage == STAGES) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer: D[128 rows x NQ] += corpus chunk * queries^T =====================
        if (lane == 0) {
            constexpr uint32_t idesc = make_idesc_bf16(TILE_ROWS_CTA, N_MMA);
            mbar_wait(bar_qfull, 0);
            tc_fence_after();
            int stage = 0;
            uint32_t phase = 0;
            uint32_t it = 0;
            for (int64_t t = cta; t < num_tiles; t += ncta, ++it) {
                const uint32_t buf = it % S_TMEM_BUFS;
                const uint32_t bphase = (it / S_TMEM_BUFS) & 1;
                mbar_wait(bar_tempty + 8 * buf, bphase ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + buf * S_ACC_STRIDE;
                // K-chunks in pairs per barrier round trip (see scan_mma.cu); an odd width ends with a single chunk
#pragma unroll 1
                for (int kc = 0; kc < k_chunks; kc += 2) {
                    const int n_here = (kc + 1 < k_chunks) ? 2 : 1;
                    const int s0 = stage;
                    const uint32_t p0 = phase;
                    int s1 = stage + 1;
                    uint32_t p1 = phase;
                    if (s1 == STAGES) {
                        s1 = 0;
                        p1 ^= 1;
                    }
                    mbar_wait(bar_full + 8 * s0, p0);
                    if (n_here == 2) mbar_wait(bar_full + 8 * s1, p1);
                    tc_fence_after();
                    for (int h = 0; h < n_here; ++h) {
                        const int sh = h == 0 ? s0 : s1;
                        const uint32_t a_addr = smem_u32(smem_ring + sh * STAGE_BYTES);
                        const uint32_t b_addr = smem_u32(smem_q + (kc + h) * Plan::Q_CHUNK);
#pragma unroll
                        for (int k4 = 0; k4 < K_CHUNK / UMMA_K; ++k4) {
                            const uint64_t adesc = make_kmajor_sw128_desc(a_addr + k4 * UMMA_K * 2);
                            const uint64_t bdesc = make_kmajor_sw128_desc(b_addr + k4 * UMMA_K * 2);
                            tc_mma_bf16<1>(d_tmem, adesc, bdesc, idesc, ((kc + h) | k4) != 0 ? 1u : 0u);
                        }
                        tc_commit<1>(bar_empty + 8 * sh);
                    }
                    if (n_here == 2) {
                        stage = s1;
                        phase = p1;
                    }
                    if (++stage == STAGES) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
                tc_commit<1>(bar_tfull + 8 * buf);
            }
        }
    } else {
        // ===================== epilogue: one TMEM lane = one corpus row =====================
        const int quarter = warp & 3;  // TMEM lane quarter this warp may access = rows [32 quarter, +32) of the tile
        const int ew = warp - 2;       // epilogue warp index; warps ew and ew + 4 share a quarter and split the queries
        const int qbase = (ew >> 2) * NQH;  // first query (of the launch's block) this warp serves
        const int nqw = max(0, min(NQH, nq - qbase));  // live queries of this warp
        // [NQH][CAP], sorted descending: this warp's own lists, or (LM_SHARED) the CTA's lists of the queries it serves
        uint64_t *my_lists = LMODE == LM_SHARED ? lists + static_cast<size_t>(qbase) * CAP
                                                : lists + static_cast<size_t>(ew) * NQH * CAP;
        int *my_locks = locks + qbase;
        float *my_stash = reinterpret_cast<float *>(smem + plan.stash_off) + static_cast<size_t>(ew) * NQH * 32;
        float tau[NQH];
#pragma unroll
        for (int q = 0; q < NQH; ++q) tau[q] = q < nqw ? -INFINITY : INFINITY;  // padded queries never pass
        // shared thresholds (see scan_mma.cu): this CTA raises slot cta % k' of a query to the best score it holds
        // With fewer CTAs than slots (k' = 256 on 148 SMs) CTA c owns the slots c, c + P, c + 2P, ... and raises slot
        // c + r P to the r-th best score it holds: every slot is still backed by a row of its own, so the minimum over
        // the k' slots stays a score that k' distinct rows reach.  (Without this the slots beyond P stayed empty, the
        // shared threshold never switched on and a top-100 search of 16 queries took 7 x the corpus read,
        // profiles/r02_k2s_wide_lists_10m.jsonl.)
        uint32_t *my_slots = tau_g + static_cast<size_t>(q0 + qbase) * ksel + (cta % ksel);  // + q * ksel
        const int n_own = ncta >= ksel ? 1 : min(32, (ksel - cta + ncta - 1) / ncta);
        // histogram refresh: a half-warp per query, two queries per pass; cb_prev keeps, 4 bits per pass, the coarse bin
        // this half-warp's query had its threshold in at the last refresh (its fine counters are fetched speculatively)
        const int half = lane >> 4, hl = lane & 15;
        uint32_t cb_prev = 0u;
        // The FIRST tile is not inserted at once (hist_g only): with empty lists and no threshold every one of its
        // 32 x NQH scores would go through a sort + fold per query and warp -- 50 us per CTA at 64 queries, and again for
        // the second tile, whose thresholds come from one tile's worth of rows.  Instead the tile is parked in the stash,
        // a quarter of its rows are counted into the histogram, and it is replayed in front of the second tile, after a
        // refresh that by then sees the first tiles of all the CTAs: a handful of rows pass instead of all.
        constexpr uint32_t FIRST_SAMPLE = 0x11111111u;  // lanes whose rows are counted when the first tile is parked
        const bool defer_first = hist_g != nullptr;
        bool deferred = false;
        uint32_t deferred_row0 = 0u;
        uint32_t it = 0;
        uint32_t next_refresh = defer_first ? 1u : 0u;  // (nothing to read before anybody has counted anything)
        for (int64_t t = cta;; t += ncta, ++it) {
            const bool have_tile = t < num_tiles;
            if (!have_tile && !deferred) break;  // (one extra turn replays a parked tile that had no successor)
            const uint32_t buf = it % S_TMEM_BUFS;
            const uint32_t bphase = (it / S_TMEM_BUFS) & 1;
            // refresh from the other CTAs.  tau_g is laid out [query][k' slots] here, so the k' slots of a query
            // are one or two coalesced 128-byte reads for the warp (lane j reads slot j) and a warp min: one L2
            // request per query instead of k' (the requests of all CTAs meet on the same few lines).
            // (Tried: a seventh warp polling the slots every 1.5 us and handing the thresholds over through shared
            // memory -- 5-8 % slower at every batch size: the polling traffic costs more than the refresh.)
            // (schedule as in K2: every tile at first, then geometrically thinning out to every 64th tile)
            const bool refresh_now = have_tile && it >= next_refresh && !(dbg & 64);
            if (refresh_now) next_refresh = it + 1u + min(it >> 1, 63u);
            // the refresh in front of the replay first waits for this tile's scores: time for the other CTAs' counts to land
            if (refresh_now && deferred) mbar_wait(bar_tfull + 8 * buf, bphase);
            if (refresh_now) {
                // The four warps of a group (one per lane quarter, same queries) split the work: each refreshes a quarter
                // of the group's queries and leaves the result in shared memory; after a barrier of the group every warp
                // picks up all of them.  (Each warp refreshing all its queries took 12 us a time at 64 queries -- five
                // tile periods, more than the four accumulators buffer.)
                constexpr int RQ = NQH / 4;
                const int rq0 = quarter * RQ;  // this warp refreshes the group's queries [rq0, rq0 + RQ)
                float *tau_sh = tau_s + qbase;
                uint32_t x[RQ];
#pragma unroll
                for (int i = 0; i < RQ; ++i) {
                    x[i] = 0u;
                    if (rq0 + i < nqw) {  // warp-uniform
                        const uint32_t *sp = tau_g + static_cast<size_t>(q0 + qbase + rq0 + i) * ksel + lane;
                        asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(x[i]) : "l"(sp));
#pragma unroll
                        for (int j = 1; j < KPL; ++j) {  // k' = 32 KPL slots: lane l reads slots l, l + 32, ...
                            uint32_t y;
                            asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(y) : "l"(sp + 32 * j));
                            x[i] = min(x[i], y);
                        }
                    }
                }
                float nv[RQ];
#pragma unroll
                for (int i = 0; i < RQ; ++i) {
                    const uint32_t m = __reduce_min_sync(FULL_MASK, x[i]);
                    nv[i] = m != 0u ? unorder_bits(m) : -INFINITY;
                }
                if (hist_g != nullptr) {
                    // Per query: the 16 coarse counters give the coarse bin the k'-th best counted row lies in, that bin's 16
                    // fine counters the fine bin; its lower edge is the threshold.  A half-warp per query, two queries per
                    // pass.  The fine counters are fetched along with the coarse ones for the bin of the last refresh
                    // (one round trip); only when a query has moved to another coarse bin is there a second round.
                    // Counters only grow and each counts distinct rows already scanned, so a stale or half-updated
                    // view only loosens the threshold.
#pragma unroll 1
                    for (int round = 0; round < 2; ++round) {
                        uint32_t cs[RQ / 2], cf[RQ / 2];
#pragma unroll
                        for (int p = 0; p < RQ / 2; ++p) {
                            cs[p] = 0u;
                            cf[p] = 0u;
                            if (rq0 + 2 * p + half < nqw) {
                                const uint32_t *hp = hist_g + static_cast<size_t>(q0 + qbase + rq0 + 2 * p + half) * SCORE_HIST_WORDS;
                                const int cbp = static_cast<int>(cb_prev >> (4 * p)) & 15;
                                asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(cs[p]) : "l"(hp + hl));
                                asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(cf[p]) : "l"(hp + 16 + 16 * cbp + hl));
                            }
                        }
                        bool moved = false;
#pragma unroll
                        for (int p = 0; p < RQ / 2; ++p) {
                            uint32_t sc = cs[p], sf = cf[p];  // -> rows counted in the coarse / fine bins >= hl
#pragma unroll
                            for (int d = 1; d < 16; d <<= 1) {
                                const uint32_t oc = __shfl_down_sync(FULL_MASK, sc, d, 16);
                                const uint32_t of = __shfl_down_sync(FULL_MASK, sf, d, 16);
                                if (hl + d < 16) {
                                    sc += oc;
                                    sf += of;
                                }
                            }
                            const unsigned bc = (__ballot_sync(FULL_MASK, sc >= static_cast<uint32_t>(ksel)) >> (16 * half)) & 0xffffu;
                            const int cb = bc != 0u ? 31 - __clz(bc) : -1;  // -1: fewer than k' rows counted so far
                            const uint32_t nxt = __shfl_sync(FULL_MASK, sc, (half << 4) | ((cb + 1) & 15));
                            const uint32_t above = (cb >= 0 && cb < 15) ? nxt : 0u;  // rows in the coarse bins above cb
                            const int cbp = static_cast<int>(cb_prev >> (4 * p)) & 15;
                            const bool same = cb == cbp;
                            const unsigned bf = (__ballot_sync(FULL_MASK, same && above + sf >= static_cast<uint32_t>(ksel)) >> (16 * half)) & 0xffffu;
                            // (no fine bin reaches k': the fine counters lag the coarse one, or belong to another bin --
                            //  the coarse edge still holds)
                            const int bin = cb < 0 ? 0 : 16 * cb + (bf != 0u ? 31 - __clz(bf) : 0);
                            if (cb >= 0 && !same) {
                                cb_prev = (cb_prev & ~(15u << (4 * p))) | (static_cast<uint32_t>(cb) << (4 * p));
                                moved = true;
                            }
                            const float edge = bin > 0 ? static_cast<float>(bin) * (1.0f / 256.0f) : -INFINITY;
                            nv[2 * p] = fmaxf(nv[2 * p], __shfl_sync(FULL_MASK, edge, 0));
                            nv[2 * p + 1] = fmaxf(nv[2 * p + 1], __shfl_sync(FULL_MASK, edge, 16));
                        }
                        if (!__any_sync(FULL_MASK, moved)) break;
                    }
                }
                if (lane == 0) {
#pragma unroll
                    for (int i = 0; i < RQ; ++i) tau_sh[rq0 + i] = fmaxf(tau_sh[rq0 + i], nv[i]);  // (this warp is the only writer)
                }
                asm volatile("bar.sync %0, 128;" ::"r"(2 + (ew >> 2)) : "memory");  // the group's four warps
#pragma unroll
                for (int q = 0; q < NQH; ++q) tau[q] = fmaxf(tau[q], tau_sh[q]);  // (padded queries stay at +inf)
            }
            float cn = 0.0f;  // l2: this thread's row norm, requested before the accumulator wait hides its latency
            if (norm2 != nullptr && have_tile) {
                const int64_t nrow = t * TILE_ROWS_CTA + quarter * 32 + lane;
                if (nrow < n_rows) cn = __ldg(norm2 + nrow);
            }
            if (have_tile) {
                mbar_wait(bar_tfull + 8 * buf, bphase);
                tc_fence_after();
            }
            // pass 0 replays the parked first tile (scores from the stash), pass 1 takes this tile's from the accumulator
#pragma unroll 1
            for (int pass = deferred ? 0 : 1; pass < (have_tile ? 2 : 1); ++pass) {
            float v[NQH];
            uint32_t row0;
            uint32_t count_lanes;  // lanes whose passing rows are still to be counted into the histogram
            if (pass == 0) {
#pragma unroll
                for (int q = 0; q < NQH; ++q) v[q] = my_stash[q * 32 + lane];
                __syncwarp();
                row0 = deferred_row0;
                count_lanes = ~FIRST_SAMPLE;
                deferred = false;
            } else {
                uint32_t r[NQH * (1 + SPLIT)];  // this warp's queries: hi columns [qbase, +NQH), lo columns NQ further on
                const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + buf * S_ACC_STRIDE + qbase;
#pragma unroll
                for (int c = 0; c < NQH / 16; ++c) tmem_ld16_nowait(taddr + c * 16, r + c * 16);
                if (SPLIT) {
#pragma unroll
                    for (int c = 0; c < NQH / 16; ++c) tmem_ld16_nowait(taddr + NQ + c * 16, r + NQH + c * 16);
                }
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                // the scores are in registers: hand the accumulator back before looking at them
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(bar_tempty + 8 * buf);
#pragma unroll
                for (int q = 0; q < NQH; ++q) v[q] = SPLIT ? __uint_as_float(r[q]) + __uint_as_float(r[NQH + q]) : __uint_as_float(r[q]);
                if (norm2 != nullptr) {  // warp-uniform
#pragma unroll
                    for (int q = 0; q < NQH; ++q) v[q] = fmaf(2.0f, v[q], -cn);
                }
                row0 = static_cast<uint32_t>(t * TILE_ROWS_CTA);
                count_lanes = 0xffffffffu;
            }
            if (it == 0u && pass == 1 && defer_first) {
                // park the first tile: scores to the stash, a quarter of the rows into the histogram
                const uint32_t row = row0 + quarter * 32 + lane;
                bool valid = row < n_rows;
                if (valid && keys_or_null != nullptr) valid = keys_or_null[row] != KEY_TOMBSTONE;
#pragma unroll
                for (int q = 0; q < NQH; ++q) my_stash[q * 32 + lane] = v[q];
                if (valid && ((FIRST_SAMPLE >> lane) & 1u))
                    for (int q = 0; q < nqw; ++q)
                        hist_count(hist_g + static_cast<size_t>(q0 + qbase + q) * SCORE_HIST_WORDS, my_stash[q * 32 + lane]);
                __syncwarp();
                deferred = true;
                deferred_row0 = row0;
                continue;
            }
            if (it == 0u && LMODE != LM_SHARED && !defer_first) {
                // first tile of this warp: every list is empty and every row would pass one by one (32 x NQ sorted
                // inserts).  Load the lists in bulk instead: per query one 32-key bitonic sort of the tile's scores.
                const uint32_t row = row0 + quarter * 32 + lane;
                bool valid = row < n_rows;
                if (valid && keys_or_null != nullptr) valid = keys_or_null[row] != KEY_TOMBSTONE;
#pragma unroll
                for (int q = 0; q < NQH; ++q) my_stash[q * 32 + lane] = v[q];
                __syncwarp();
                for (int q = 0; q < nqw; ++q) {
                    uint64_t key = valid ? pack_key(my_stash[q * 32 + lane], row) : 0ull;
                    key = bitonic_sort32_desc(key, lane);
                    my_lists[q * CAP + lane] = key;  // entries 32 .. CAP-1 stay empty
                    if (KPL == 1) {  // a full list already gates
                        const uint32_t last = __shfl_sync(FULL_MASK, static_cast<uint32_t>(key >> 32), 31);
                        const float nt = last != 0u ? unorder_bits(last) : -INFINITY;
                        switch (q) {
#define FR_TAU_CASE(i) case i: tau[(i) < NQH ? (i) : 0] = fmaxf(tau[(i) < NQH ? (i) : 0], nt); break;
#define FR_TAU_CASE8(b) FR_TAU_CASE(b) FR_TAU_CASE(b + 1) FR_TAU_CASE(b + 2) FR_TAU_CASE(b + 3) \
                        FR_TAU_CASE(b + 4) FR_TAU_CASE(b + 5) FR_TAU_CASE(b + 6) FR_TAU_CASE(b + 7)
                            FR_TAU_CASE8(0) FR_TAU_CASE8(8)
                            FR_TAU_CASE8(16) FR_TAU_CASE8(24)
#undef FR_TAU_CASE8
#undef FR_TAU_CASE
                            default: break;
                        }
                    }
                    const uint32_t mine_hi = static_cast<uint32_t>(key >> 32);  // lane r: the r-th best row of this warp
                    if (lane < n_own && mine_hi != 0u) atomicMax(my_slots + static_cast<size_t>(q) * ksel + lane * ncta, mine_hi);
                }
                __syncwarp();
                continue;
            }
            // gate: per group of 16 queries, max_q (score_q - tau_q); two chains per group
            constexpr int NG = NQH / 16;
            float gm[NG];
#pragma unroll
            for (int g = 0; g < NG; ++g) {
                float m0 = v[g * 16] - tau[g * 16], m1 = v[g * 16 + 1] - tau[g * 16 + 1];
#pragma unroll
                for (int i = 2; i < 16; i += 2) {
                    m0 = fmaxf(m0, v[g * 16 + i] - tau[g * 16 + i]);
                    m1 = fmaxf(m1, v[g * 16 + i + 1] - tau[g * 16 + i + 1]);
                }
                gm[g] = fmaxf(m0, m1);
            }
            float mall = gm[0];
#pragma unroll
            for (int g = 1; g < NG; ++g) mall = fmaxf(mall, gm[g]);
            if (!__any_sync(FULL_MASK, mall > 0.0f) || (dbg & 2)) continue;  // the common case
            // ---- candidate path (one copy of the insert code whatever NQ: the queries with a passing row are
            //      walked with a run-time index, so the scores of the groups that hold one go through shared memory) ----
            const uint32_t row = row0 + quarter * 32 + lane;
            bool valid = row < n_rows;  // rows past the end arrive as zeros from TMA
            if (valid && keys_or_null != nullptr) valid = keys_or_null[row] != KEY_TOMBSTONE;
            static_assert(NQH <= 32, "one mask word per epilogue warp");
            uint32_t pm[1] = {0u};  // this row's passing queries
#pragma unroll
            for (int g = 0; g < NG; ++g) {
                if (!__any_sync(FULL_MASK, gm[g] > 0.0f)) continue;  // warp-uniform: no row of the warp passes here
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    const int q = g * 16 + i;
                    my_stash[q * 32 + lane] = v[q];
                    pm[q / 32] |= (valid && v[q] > tau[q]) ? (1u << (q & 31)) : 0u;
                }
            }
            __syncwarp();
            const bool count_me = hist_g != nullptr && ((count_lanes >> lane) & 1u) != 0u;
#pragma unroll
            for (int h = 0; h < 1; ++h) {
                uint32_t qm = __reduce_or_sync(FULL_MASK, pm[h]);  // queries with at least one passing row
                // LM_SHARED: the four warps that share the lists start their walk a quarter of the queries apart; walking in
                // the same order they met at every lock of the early, busy tiles (14 spins per acquisition measured)
                constexpr uint32_t QMASK = NQH >= 32 ? 0xffffffffu : ((1u << (NQH & 31)) - 1u);
                const int rot = LMODE == LM_SHARED ? quarter * (NQH / 4) : 0;
                qm = ((qm >> rot) | (qm << ((NQH - rot) & 31))) & QMASK;
                while (qm != 0u) {
                    const int qb = (__ffs(qm) - 1 + rot) & (NQH - 1);
                    qm &= qm - 1;
                    const int q = h * 32 + qb;
                    const bool has = ((pm[h] >> qb) & 1u) != 0u;
                    if (count_me && has) hist_count(hist_g + static_cast<size_t>(q0 + qbase + q) * SCORE_HIST_WORDS, my_stash[q * 32 + lane]);
                    unsigned mask = __ballot_sync(FULL_MASK, has);
                    uint64_t *lp = my_lists + q * CAP;
                    if (LMODE == LM_SHARED) {  // the other lane quarters' warps fold into the same list
                        if (lane == 0)
                            while (atomicCAS(my_locks + q, 0, 1) != 0)
                                while (*(volatile int *)(my_locks + q) != 0) {}  // held for a few hundred cycles: poll, no sleep
                        __syncwarp();
                        __threadfence_block();
                    }
                    WarpTopK<KPL> lst;
#pragma unroll
                    for (int j = 0; j < KPL; ++j) lst.e[j] = LMODE == LM_SHARED ? *(volatile uint64_t *)(lp + j * 32 + lane) : lp[j * 32 + lane];
                    const uint64_t mine = pack_key(my_stash[q * 32 + lane], row);
                    if (__popc(mask) > 3) {
                        // many rows at once (before the thresholds bite): one sort + one fold network
                        const uint64_t p = bitonic_sort32_desc(has ? mine : 0ull, lane);
                        fold_sorted32<KPL>(lst.e, p, lane);
                    } else {
                        while (mask) {
                            const int src = __ffs(mask) - 1;
                            mask &= mask - 1;
                            const uint32_t lo = __shfl_sync(FULL_MASK, static_cast<uint32_t>(mine), src);
                            const uint32_t hi = __shfl_sync(FULL_MASK, static_cast<uint32_t>(mine >> 32), src);
                            lst.insert((static_cast<uint64_t>(hi) << 32) | lo, CAP, lane);
                        }
                    }
#pragma unroll
                    for (int j = 0; j < KPL; ++j) lp[j * 32 + lane] = lst.e[j];
                    if (LMODE == LM_SHARED) {
                        __threadfence_block();
                        __syncwarp();
                        if (lane == 0) atomicExch(my_locks + q, 0);
                    }
                    const float nt = key_threshold(lst.kth(CAP));
                    // tau lives in registers: a jump table of NQ one-line cases instead of NQ predicated updates
                    switch (q) {
#define FR_TAU_CASE(i) case i: tau[(i) < NQH ? (i) : 0] = fmaxf(tau[(i) < NQH ? (i) : 0], nt); break;
#define FR_TAU_CASE8(b) FR_TAU_CASE(b) FR_TAU_CASE(b + 1) FR_TAU_CASE(b + 2) FR_TAU_CASE(b + 3) \
                        FR_TAU_CASE(b + 4) FR_TAU_CASE(b + 5) FR_TAU_CASE(b + 6) FR_TAU_CASE(b + 7)
                        FR_TAU_CASE8(0) FR_TAU_CASE8(8)
                        FR_TAU_CASE8(16) FR_TAU_CASE8(24)
#undef FR_TAU_CASE8
#undef FR_TAU_CASE
                        default: break;
                    }
                    const uint32_t own_hi = static_cast<uint32_t>(lst.e[0] >> 32);  // lane r: the r-th best row of the list
                    if (lane < n_own && own_hi != 0u) atomicMax(my_slots + static_cast<size_t>(q) * ksel + lane * ncta, own_hi);
                }
            }
            __syncwarp();
            }  // pass
        }
        // ---- merge the four warps' lists of each query, write the CTA's partial lists ----
        asm volatile("bar.sync 1, %0;" ::"n"(32 * EW) : "memory");  // epilogue warps only
        for (int q = ew; q < nq; q += EW) {
            const int first = (q / NQH) * 4, ql = q % NQH;  // the four warps (one per lane quarter) that served query q
            WarpTopK<KPL> lst;
            if (LMODE == LM_SHARED) {  // already one list per query
#pragma unroll
                for (int j = 0; j < KPL; ++j) lst.e[j] = lists[static_cast<size_t>(q) * CAP + j * 32 + lane];
            } else {
#pragma unroll
                for (int j = 0; j < KPL; ++j) lst.e[j] = lists[(static_cast<size_t>(first) * NQH + ql) * CAP + j * 32 + lane];
                for (int w = 1; w < 4; ++w) lst.merge_sorted(lists + (static_cast<size_t>(first + w) * NQH + ql) * CAP, CAP, CAP, lane);
            }
            uint64_t *dst = partials + (static_cast<size_t>(cta) * nq_total + q0 + q) * ksel;
#pragma unroll
            for (int j = 0; j < KPL; ++j) dst[j * 32 + lane] = lst.e[j];
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(S_TMEM_COLS) : "memory");
    }
}

}  // namespace mma

namespace {
template <int NQ, int KPL, int SPLIT>
cudaError_t launch_small(const MmaScanArgs &a, const CUtensorMap &tq, const CUtensorMap &tc, int k_chunks, int nq, int q0,
                         size_t *alloc_out) {
    if constexpr (mma::small_list_mode(NQ, KPL) == mma::LM_NONE) {
        if (alloc_out) *alloc_out = 0;
        return cudaErrorInvalidValue;
    } else {
    const mma::SmallPlan<NQ, KPL, SPLIT> plan(k_chunks);
    if (alloc_out) {  // planning only: does this instance fit, and with how deep a ring?
        *alloc_out = plan.stages >= mma::S_MIN_STAGES ? plan.alloc : 0;
        return cudaSuccess;
    }
    if (plan.stages < mma::S_MIN_STAGES) return cudaErrorInvalidValue;
    auto kern = mma::scan_mma_small_kernel<NQ, KPL, SPLIT>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(plan.alloc));
    if (e != cudaSuccess) return e;
    kern<<<a.plan.lists, mma::SmallPlan<NQ, KPL, SPLIT>::THREADS, plan.alloc, a.stream>>>(tq, tc, a.keys_or_null, a.n_rows, nq, a.ksel, a.partials,
                                                                a.nq_total, a.tau_g, k_chunks, q0, a.nq_dev, a.norm2, a.dbg, a.hist);
    count_launch();
    return cudaGetLastError();
    }
}

cudaError_t dispatch_small(const MmaScanArgs &a, const CUtensorMap &tq, const CUtensorMap &tc, int nq_pad, int ksel,
                           int k_chunks, int split, int nq, int q0, size_t *alloc_out) {
#define FR_SMALL(NQ, KPL, SPLIT) launch_small<NQ, KPL, SPLIT>(a, tq, tc, k_chunks, nq, q0, alloc_out)
#define FR_SMALL_K(NQ, SPLIT)                                                                      \
    return ksel <= 32 ? FR_SMALL(NQ, 1, SPLIT)                                                     \
                      : (ksel <= 64 ? FR_SMALL(NQ, 2, SPLIT) : (ksel <= 128 ? FR_SMALL(NQ, 4, SPLIT) : FR_SMALL(NQ, 8, SPLIT)))
    if (ksel > 256) return cudaErrorInvalidValue;
    if (split) {
        switch (nq_pad) {
            case 16: FR_SMALL_K(16, 1);
            case 32: FR_SMALL_K(32, 1);
            default: return cudaErrorInvalidValue;  // 2 x 64 operand rows exceed an accumulator
        }
    }
    switch (nq_pad) {
        case 16: FR_SMALL_K(16, 0);
        case 32: FR_SMALL_K(32, 0);
        case 64: FR_SMALL_K(64, 0);
        default: return cudaErrorInvalidValue;
    }
#undef FR_SMALL_K
#undef FR_SMALL
}
}  // namespace

// Padded query count (16 / 32 / 64) of the instance that serves `nq` queries of width `dim` with k' = ksel
// (`split`: bf16 hi + lo halves of every query, twice the query block), 0 when none fits (too many queries for
// the shared memory left beside a ring of at least S_MIN_STAGES stages).
int scan_mma_small_nq(int nq, int ksel, int dim, int split) {
    if (nq < 1 || nq > (split ? 32 : 64) || ksel > 256 || dim < 64 || dim % 64 != 0 || dim > 1024) return 0;
    const int nq_pad = nq <= 16 ? 16 : (nq <= 32 ? 32 : 64);
    size_t alloc = 0;
    MmaScanArgs dummy{};
    CUtensorMap none{};
    if (dispatch_small(dummy, none, none, nq_pad, ksel, dim / mma::K_CHUNK, split, nq, 0, &alloc) != cudaSuccess || alloc == 0)
        return 0;
    return nq_pad;
}

// largest batch one K2s launch can serve for this width, k' and query precision (0 = none)
int scan_mma_small_max_batch(int ksel, int dim, int split) {
    for (int nq : {64, 32, 16})
        if (scan_mma_small_nq(nq, ksel, dim, split) != 0) return nq;
    return 0;
}

// One launch: queries [q0, q0 + nq) of the query block (a.nq_dev: of the *a.nq_dev live ones).  a.queries_bf16 is
// [a.nq_pad rows][dim * (1 + a.split)] bf16.
cudaError_t launch_scan_mma_small(const MmaScanArgs &a, int q0, int nq) {
    const int nq_pad = scan_mma_small_nq(nq, a.ksel, a.dim, a.split);
    if (nq_pad == 0 || a.nq_pad < q0 + nq_pad) return cudaErrorInvalidValue;
    CUtensorMap tq, tc;
    if (!mma::make_row_major_map(&tq, a.queries_bf16, a.nq_pad, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, nq_pad, a.dim * (1 + a.split)) ||
        !mma::make_row_major_map(&tc, a.corpus, a.n_rows, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, mma::TILE_ROWS_CTA, a.dim))
        return cudaErrorNotSupported;
    return dispatch_small(a, tq, tc, nq_pad, a.ksel, a.dim / mma::K_CHUNK, a.split, nq, q0, nullptr);
}

}  // namespace fr
