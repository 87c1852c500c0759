// K2 scan_topk_mma -- tensor-core exact scan with a fused top-k epilogue (sm_100a: tcgen05 + TMEM + TMA).
//
// Replaces the arithmetic behind chromadb Collection.query (parent_child/chroma_child_store.py:63,
// parent_child/multivector_store.py:151) for query batches too large for the CUDA-core stream
// kernel (K1 is FFMA-bound above ~2 queries per pass).
//
// The scan of a query block against the corpus IS a dense contraction, so it runs on the 5th-gen
// tensor cores:   D[128 queries x 128 rows] += A[128 x 16] * B[128 x 16]^T   (bf16 in, fp32 in TMEM)
//   * A = the query block (bf16 copy of the normalised queries, zero-padded to 128 rows), loaded once
//     by TMA into shared memory as six K-chunks of [128 rows x 64 elements], SWIZZLE_128B, K-major;
//   * B = corpus tiles of 128 rows, streamed by TMA ([64 x 128] boxes of the row-major bf16 corpus
//     -- its natural layout is already the K-major operand) through a 5-stage mbarrier ring;
//   * one elected thread issues 24 tcgen05.mma (K = 16 each) per corpus tile into one of four
//     128-column TMEM accumulators; tcgen05.commit frees the smem stage / publishes the accumulator;
//   * one TMEM lane = one query: each of the 128 epilogue threads reads ITS query's scores with
//     tcgen05.ld (32 columns at a time) and keeps a running maximum; only when some lane's maximum
//     beats its current threshold does the warp enter the insert path (transpose 16 columns through
//     shared memory, warp-cooperative sorted insert into that query's list in shared memory).
//     No score ever goes to HBM.
// Roofline: at <= 128 queries per pass the MMA work per tile (24 x 64 cycles) is far below the HBM
// time of the tile (128 rows x 768 B), so this kernel is HBM-bound like K1, but for 128 queries at once.
//
// Exactness.  The tensor cores need bf16 queries; the reference semantics (and K1) use fp32 queries.
// So K2 only SELECTS: it keeps k' = 32*KPL >= 2k candidates per query, ranked by the bf16-query
// score; rescore_kernel then recomputes the k' scores with the fp32 query (fp32 FMA), sorts them
// by the exact key and CERTIFIES the top-k:  every row outside the candidate set has
//     exact score <= (k'-th selection score) + |q - bf16(q)|_2 * |c|_2  (+ fp32 accumulation slack),
// so if the k-th exact score is above that bound the result equals the fp32-query scan; otherwise
// the query is flagged and re-scanned by the stream kernel (scan_stream_fallback_kernel).
#include <cuda.h>

#include "fr_common.cuh"
#include "fr_kernels.h"

namespace fr {
namespace mma {

constexpr int DIM = 384;
constexpr int M_TILE = 128;                    // queries per pass (TMEM lanes)
constexpr int N_TILE = 128;                    // corpus rows per accumulator
constexpr int K_CHUNK = 64;                    // bf16 elements per 128-byte swizzle row
constexpr int K_CHUNKS = DIM / K_CHUNK;        // 6
constexpr int UMMA_K = 16;
constexpr int CHUNK_BYTES = N_TILE * K_CHUNK * 2;  // 16 KB: one [128 x 64] bf16 box (A chunk or B stage)
constexpr int TMEM_BUFS = 4;
constexpr int TMEM_COLS = TMEM_BUFS * N_TILE;  // 512
constexpr int THREADS = 192;                   // warp 0 TMA, warp 1 MMA + TMEM alloc, warps 2-5 epilogue
constexpr int TR_STRIDE = 17;                  // transpose scratch row stride (floats), conflict-free

template <int KPL>
struct SmemPlan {
    static constexpr int STAGES = (KPL == 1) ? 5 : 3;
    static constexpr int CAP = 32 * KPL;
    static constexpr size_t A_OFF = 0;
    static constexpr size_t B_OFF = A_OFF + size_t(K_CHUNKS) * CHUNK_BYTES;
    static constexpr size_t LIST_OFF = B_OFF + size_t(STAGES) * CHUNK_BYTES;
    static constexpr size_t TR_OFF = LIST_OFF + size_t(M_TILE) * CAP * 8;
    static constexpr size_t BAR_OFF = TR_OFF + size_t(4) * 32 * TR_STRIDE * 4;
    static constexpr size_t TOTAL = BAR_OFF + 256;
    static constexpr size_t ALLOC = TOTAL + 1024;  // slack to align the base to 1024 B
};

// ---- PTX wrappers -----------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra DONE;\n"
        "bra LAB_WAIT;\n"
        "DONE:\n"
        "}" ::"r"(bar),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap *map, uint32_t bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
        "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_mma_bf16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
        "}" ::"r"(d_tmem),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum)
        : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// Shared-memory matrix descriptor: K-major, SWIZZLE_128B, rows 128 B apart, 8-row groups 1024 B apart
// (cute::UMMA::SmemDescriptor: start>>4 [0,14), LBO>>4 [16,30), SBO>>4 [32,46), version=1 [46,48),
//  layout_type=2 (SWIZZLE_128B) [61,64)).
__device__ __forceinline__ uint64_t make_kmajor_sw128_desc(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= static_cast<uint64_t>((smem_addr & 0x3ffffu) >> 4);
    d |= static_cast<uint64_t>(1) << 16;             // leading byte offset (unused for swizzled K-major)
    d |= static_cast<uint64_t>(1024 >> 4) << 32;     // stride byte offset: 8 rows x 128 B
    d |= static_cast<uint64_t>(1) << 46;             // descriptor version (Blackwell)
    d |= static_cast<uint64_t>(2) << 61;             // SWIZZLE_128B
    return d;
}
// Instruction descriptor (cute::UMMA::InstrDescriptor): D=F32 [4,6), A=BF16 [7,10), B=BF16 [10,13),
// A,B K-major (bits 15,16 = 0), N>>3 [17,23), M>>4 [24,29).
__host__ __device__ constexpr uint32_t make_idesc_bf16(int m, int n) {
    return (1u << 4) | (1u << 7) | (1u << 10) | (static_cast<uint32_t>(n >> 3) << 17) |
           (static_cast<uint32_t>(m >> 4) << 24);
}

// ---------------------------------------------------------------------------------------------
template <int KPL>
__global__ void __launch_bounds__(THREADS, 1)
scan_mma_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_c,
                const int64_t *__restrict__ keys_or_null, int64_t n_rows, int nq, int q_row0, int ksel,
                uint64_t *__restrict__ partials /* [gridDim.x][nq_total][ksel] */, int nq_total, int q_offset) {
    using Plan = SmemPlan<KPL>;
    constexpr int STAGES = Plan::STAGES;
    constexpr int CAP = Plan::CAP;
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t *smem_a = smem + Plan::A_OFF;
    uint8_t *smem_b = smem + Plan::B_OFF;
    uint64_t *lists = reinterpret_cast<uint64_t *>(smem + Plan::LIST_OFF);
    float *tr_all = reinterpret_cast<float *>(smem + Plan::TR_OFF);
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + Plan::BAR_OFF);
    // barrier slots: full[STAGES] | empty[STAGES] | tmem_full[4] | tmem_empty[4] | a_full | tmem_ptr
    const uint32_t bar_full = smem_u32(bars);
    const uint32_t bar_empty = smem_u32(bars + STAGES);
    const uint32_t bar_tfull = smem_u32(bars + 2 * STAGES);
    const uint32_t bar_tempty = smem_u32(bars + 2 * STAGES + TMEM_BUFS);
    const uint32_t bar_afull = smem_u32(bars + 2 * STAGES + 2 * TMEM_BUFS);
    uint32_t *tmem_ptr_smem = reinterpret_cast<uint32_t *>(bars + 2 * STAGES + 2 * TMEM_BUFS + 1);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(bar_full + 8 * s, 1);
            mbar_init(bar_empty + 8 * s, 1);
        }
        for (int b = 0; b < TMEM_BUFS; ++b) {
            mbar_init(bar_tfull + 8 * b, 1);
            mbar_init(bar_tempty + 8 * b, 4);  // one arrival per epilogue warp
        }
        mbar_init(bar_afull, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr_smem)),
                     "r"(TMEM_COLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    // epilogue warps clear their queries' lists while the allocation happens
    if (warp >= 2) {
        for (int i = threadIdx.x - 64; i < M_TILE * CAP; i += 128) lists[i] = 0ull;
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;

    const int64_t num_tiles = (n_rows + N_TILE - 1) / N_TILE;

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            mbar_expect_tx(bar_afull, K_CHUNKS * CHUNK_BYTES);
#pragma unroll
            for (int kc = 0; kc < K_CHUNKS; ++kc)
                tma_load_2d(smem_u32(smem_a + kc * CHUNK_BYTES), &tmap_q, bar_afull, kc * K_CHUNK, q_row0);
            int stage = 0;
            uint32_t phase = 0;
            for (int64_t t = blockIdx.x; t < num_tiles; t += gridDim.x) {
                const int row0 = static_cast<int>(t * N_TILE);
#pragma unroll 1
                for (int kc = 0; kc < K_CHUNKS; ++kc) {
                    mbar_wait(bar_empty + 8 * stage, phase ^ 1);
                    mbar_expect_tx(bar_full + 8 * stage, CHUNK_BYTES);
                    tma_load_2d(smem_u32(smem_b + stage * CHUNK_BYTES), &tmap_c, bar_full + 8 * stage, kc * K_CHUNK, row0);
                    if (++stage == STAGES) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (lane == 0) {
            constexpr uint32_t idesc = make_idesc_bf16(M_TILE, N_TILE);
            mbar_wait(bar_afull, 0);
            tc_fence_after();
            int stage = 0;
            uint32_t phase = 0;
            uint32_t it = 0;
            for (int64_t t = blockIdx.x; t < num_tiles; t += gridDim.x, ++it) {
                const uint32_t buf = it % TMEM_BUFS;
                const uint32_t bphase = (it / TMEM_BUFS) & 1;
                mbar_wait(bar_tempty + 8 * buf, bphase ^ 1);  // epilogue has drained this accumulator
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + buf * N_TILE;
#pragma unroll 1
                for (int kc = 0; kc < K_CHUNKS; ++kc) {
                    mbar_wait(bar_full + 8 * stage, phase);
                    tc_fence_after();
                    const uint32_t a_addr = smem_u32(smem_a + kc * CHUNK_BYTES);
                    const uint32_t b_addr = smem_u32(smem_b + stage * CHUNK_BYTES);
#pragma unroll
                    for (int k4 = 0; k4 < K_CHUNK / UMMA_K; ++k4) {
                        const uint64_t adesc = make_kmajor_sw128_desc(a_addr + k4 * UMMA_K * 2);
                        const uint64_t bdesc = make_kmajor_sw128_desc(b_addr + k4 * UMMA_K * 2);
                        tc_mma_bf16(d_tmem, adesc, bdesc, idesc, (kc | k4) != 0 ? 1u : 0u);
                    }
                    tc_commit(bar_empty + 8 * stage);  // smem stage reusable once these MMAs have read it
                    if (++stage == STAGES) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
                tc_commit(bar_tfull + 8 * buf);  // accumulator complete
            }
        }
    } else {
        // ===================== epilogue: one TMEM lane = one query =====================
        const int quarter = warp & 3;              // TMEM lane quarter this warp may access
        const int my_q = quarter * 32 + lane;      // query index inside the block
        float *tr = tr_all + (warp - 2) * 32 * TR_STRIDE;
        uint64_t *my_lists = lists + static_cast<size_t>(quarter) * 32 * CAP;
        float tau = (my_q < nq) ? -INFINITY : INFINITY;  // padded query rows never pass the gate
        uint32_t it = 0;
        for (int64_t t = blockIdx.x; t < num_tiles; t += gridDim.x, ++it) {
            const uint32_t buf = it % TMEM_BUFS;
            const uint32_t bphase = (it / TMEM_BUFS) & 1;
            mbar_wait(bar_tfull + 8 * buf, bphase);
            tc_fence_after();
            const uint32_t row0 = static_cast<uint32_t>(t * N_TILE);
#pragma unroll 1
            for (int c = 0; c < N_TILE / 32; ++c) {
                float v[32];
                tmem_ld32(tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + buf * N_TILE + c * 32, v);
                float mx = v[0];
#pragma unroll
                for (int i = 1; i < 32; ++i) mx = fmaxf(mx, v[i]);
                if (__ballot_sync(FULL_MASK, mx > tau) == 0u) continue;  // the common case
                // ---- insert path: 16 columns at a time through the transpose scratch ----
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    float hmx = v[h * 16];
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        tr[lane * TR_STRIDE + i] = v[h * 16 + i];
                        hmx = fmaxf(hmx, v[h * 16 + i]);
                    }
                    __syncwarp();
                    unsigned m = __ballot_sync(FULL_MASK, hmx > tau);
                    while (m) {
                        const int tq = __ffs(m) - 1;  // the lane (query) that has candidates
                        m &= m - 1;
                        const float tau_t = __shfl_sync(FULL_MASK, tau, tq);
                        const float val = (lane < 16) ? tr[tq * TR_STRIDE + lane] : -INFINITY;
                        const uint32_t row = row0 + c * 32 + h * 16 + lane;
                        unsigned cm = __ballot_sync(FULL_MASK, lane < 16 && val > tau_t && row < n_rows);
                        if (cm == 0u) continue;
                        WarpTopK<KPL> lst;
                        uint64_t *lp = my_lists + static_cast<size_t>(tq) * CAP;
#pragma unroll
                        for (int j = 0; j < KPL; ++j) lst.e[j] = lp[j * 32 + lane];
                        while (cm) {
                            const int src = __ffs(cm) - 1;
                            cm &= cm - 1;
                            const float sv = __shfl_sync(FULL_MASK, val, src);
                            const uint32_t rv = row0 + c * 32 + h * 16 + src;
                            if (keys_or_null != nullptr && keys_or_null[rv] == KEY_TOMBSTONE) continue;
                            lst.insert(pack_key(sv, rv), ksel, lane);
                        }
                        lst.store(lp, lane);
                        const float new_tau = key_threshold(lst.kth(ksel));
                        if (lane == tq) tau = new_tau;
                    }
                    __syncwarp();
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_tempty + 8 * buf);
        }
        // ---- this CTA's list for each of its queries ----
        __syncwarp();
        for (int q = 0; q < 32; ++q) {
            const int qq = quarter * 32 + q;
            if (qq >= nq) break;
            uint64_t *dst = partials + (static_cast<size_t>(blockIdx.x) * nq_total + q_offset + qq) * ksel;
            const uint64_t *lp = my_lists + static_cast<size_t>(q) * CAP;
            for (int i = lane; i < ksel; i += 32) dst[i] = lp[i];
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
    }
}

// ---------------------------------------------------------------------------------------------
// Query preparation for K2: bf16 copy (zero-padded to a multiple of 128 rows) and the per-query
// selection error bound |q - bf16(q)|_2.
__global__ void prep_queries_kernel(const float *__restrict__ q, int nq, int nq_pad, __nv_bfloat16 *__restrict__ qb,
                                    float *__restrict__ err_bound) {
    const int row = blockIdx.x;
    const int lane = threadIdx.x;  // 32 threads
    float ss = 0.0f;
    for (int e = lane; e < DIM; e += 32) {
        const float x = row < nq ? q[static_cast<size_t>(row) * DIM + e] : 0.0f;
        const __nv_bfloat16 h = __float2bfloat16_rn(x);
        qb[static_cast<size_t>(row) * DIM + e] = h;
        const float d = x - __bfloat162float(h);
        ss = fmaf(d, d, ss);
    }
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) ss += __shfl_xor_sync(FULL_MASK, ss, s);
    if (lane == 0 && row < nq) err_bound[row] = sqrtf(ss);
    (void)nq_pad;
}

// ---------------------------------------------------------------------------------------------
// rescore_kernel: exact fp32-query scores of the k' selected rows, exact order, certification.
// One CTA (4 warps) per query.  sel: [B][ksel] packed keys sorted by selection score.
template <int KPL>
__global__ void __launch_bounds__(128)
rescore_kernel(const uint64_t *__restrict__ sel, int ksel, const float *__restrict__ queries,
               const uint8_t *__restrict__ corpus, const int64_t *__restrict__ row_keys,
               const float *__restrict__ err_bound, int k, float *__restrict__ out_dist,
               uint64_t *__restrict__ out_packed, int64_t *__restrict__ out_keys, uint8_t *__restrict__ flags,
               int *__restrict__ fail_count, int *__restrict__ fail_list) {
    __shared__ float sq[DIM];
    __shared__ uint64_t exact[32 * KPL];
    const int b = blockIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int e = threadIdx.x; e < DIM; e += blockDim.x) sq[e] = queries[static_cast<size_t>(b) * DIM + e];
    for (int i = threadIdx.x; i < 32 * KPL; i += blockDim.x) exact[i] = 0ull;
    __syncthreads();
    const uint64_t *s = sel + static_cast<size_t>(b) * ksel;
    for (int j = warp; j < ksel; j += 4) {
        const uint64_t key = s[j];
        if (key == 0ull) continue;  // warp-uniform
        const uint32_t row = key_row(key);
        // lane l owns elements [12 l, 12 l + 12): three 8-byte loads
        const uint2 *rp = reinterpret_cast<const uint2 *>(corpus + static_cast<size_t>(row) * (DIM * 2) + lane * 24);
        float acc = 0.0f;
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            const uint2 w = rp[i];
            const float *qq = sq + lane * 12 + i * 4;
            acc = fmaf(__uint_as_float(w.x << 16), qq[0], acc);
            acc = fmaf(__uint_as_float(w.x & 0xffff0000u), qq[1], acc);
            acc = fmaf(__uint_as_float(w.y << 16), qq[2], acc);
            acc = fmaf(__uint_as_float(w.y & 0xffff0000u), qq[3], acc);
        }
#pragma unroll
        for (int sft = 16; sft > 0; sft >>= 1) acc += __shfl_xor_sync(FULL_MASK, acc, sft);
        if (lane == 0) exact[j] = pack_key(acc, row);
    }
    __syncthreads();
    if (warp != 0) return;
    WarpTopK<KPL> lst;
    lst.clear();
    int n_valid = 0;
    for (int j = 0; j < ksel; ++j) {
        const uint64_t key = exact[j];
        if (key == 0ull) continue;
        ++n_valid;
        lst.insert(key, 32 * KPL, lane);
    }
    // certification: anything outside the candidate set scores at most (last selection score + bound)
    bool certified = true;
    const uint64_t last_sel = s[ksel - 1];
    if (last_sel != 0ull && n_valid >= k) {
        const float cut = key_score(last_sel) + err_bound[b] * 1.004f + 4e-6f;
        const float kth_exact = key_score(lst.kth(k));
        certified = kth_exact > cut;
    }
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
            if (out_dist) out_dist[o] = 1.0f - key_score(key);
            if (out_packed) out_packed[o] = key;
            out_keys[o] = row_keys[key_row(key)];
        }
    }
    if (lane == 0) {
        flags[b] = certified ? 0 : 1;
        if (!certified) fail_list[atomicAdd(fail_count, 1)] = b;
    }
}

// ---------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (fn) return fn;
    void *p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) != cudaSuccess ||
        qres != cudaDriverEntryPointSuccess)
        return nullptr;
    fn = reinterpret_cast<EncodeTiledFn>(p);
    return fn;
}

// [rows][384] bf16 row-major -> boxes of [64 elements x 128 rows], 128-byte swizzle, OOB rows read as zero
static bool make_row_major_map(CUtensorMap *map, const void *base, int64_t rows) {
    EncodeTiledFn fn = get_encode_fn();
    if (!fn) return false;
    const cuuint64_t gdim[2] = {static_cast<cuuint64_t>(DIM), static_cast<cuuint64_t>(rows)};
    const cuuint64_t gstride[1] = {static_cast<cuuint64_t>(DIM * 2)};
    const cuuint32_t box[2] = {K_CHUNK, N_TILE};
    const cuuint32_t estr[2] = {1, 1};
    return fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void *>(base), gdim, gstride, box, estr,
              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

}  // namespace mma

// ---- host launchers ---------------------------------------------------------------------------
int scan_mma_ksel(int k) { return k <= 16 ? 32 : (k <= 32 ? 64 : 0); }

cudaError_t launch_prep_queries(const float *q, int nq, int nq_pad, void *qb, float *err_bound, cudaStream_t s) {
    mma::prep_queries_kernel<<<nq_pad, 32, 0, s>>>(q, nq, nq_pad, static_cast<__nv_bfloat16 *>(qb), err_bound);
    count_launch();
    return cudaGetLastError();
}

cudaError_t launch_scan_mma(const MmaScanArgs &a) {
    CUtensorMap tq, tc;
    if (!mma::make_row_major_map(&tq, a.queries_bf16, a.nq_pad) || !mma::make_row_major_map(&tc, a.corpus, a.n_rows))
        return cudaErrorNotSupported;
    const int ksel = a.ksel;
    for (int g = 0; g * mma::M_TILE < a.nq_total; ++g) {
        const int q0 = g * mma::M_TILE;
        const int nq = (a.nq_total - q0 < mma::M_TILE) ? (a.nq_total - q0) : mma::M_TILE;
        cudaError_t e;
        if (ksel <= 32) {
            auto kern = mma::scan_mma_kernel<1>;
            e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     static_cast<int>(mma::SmemPlan<1>::ALLOC));
            if (e != cudaSuccess) return e;
            kern<<<a.grid, mma::THREADS, mma::SmemPlan<1>::ALLOC, a.stream>>>(tq, tc, a.keys_or_null, a.n_rows, nq, q0,
                                                                              ksel, a.partials, a.nq_total, q0);
        } else {
            auto kern = mma::scan_mma_kernel<2>;
            e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     static_cast<int>(mma::SmemPlan<2>::ALLOC));
            if (e != cudaSuccess) return e;
            kern<<<a.grid, mma::THREADS, mma::SmemPlan<2>::ALLOC, a.stream>>>(tq, tc, a.keys_or_null, a.n_rows, nq, q0,
                                                                              ksel, a.partials, a.nq_total, q0);
        }
        count_launch();
        e = cudaGetLastError();
        if (e != cudaSuccess) return e;
    }
    return cudaSuccess;
}

int scan_mma_plan_grid(int sm_count, int64_t n_rows) {
    const int64_t tiles = (n_rows + mma::N_TILE - 1) / mma::N_TILE;
    if (tiles < 1) return 1;
    return static_cast<int>(tiles < sm_count ? tiles : sm_count);
}

cudaError_t launch_rescore(const RescoreArgs &a) {
    if (a.B <= 0) return cudaSuccess;
    if (a.ksel <= 32)
        mma::rescore_kernel<1><<<a.B, 128, 0, a.stream>>>(a.sel, a.ksel, a.queries, a.corpus, a.row_keys, a.err_bound,
                                                         a.k, a.out_dist, a.out_packed, a.out_keys, a.flags,
                                                         a.fail_count, a.fail_list);
    else
        mma::rescore_kernel<2><<<a.B, 128, 0, a.stream>>>(a.sel, a.ksel, a.queries, a.corpus, a.row_keys, a.err_bound,
                                                         a.k, a.out_dist, a.out_packed, a.out_keys, a.flags,
                                                         a.fail_count, a.fail_list);
    count_launch();
    return cudaGetLastError();
}

}  // namespace fr
