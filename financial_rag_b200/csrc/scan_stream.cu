// K1 scan_topk_stream -- HBM-streaming exact scan with a fused register top-k (sm_100a).
//
// Replaces the arithmetic behind chromadb Collection.query as called from
// parent_child/chroma_child_store.py:63 for small query batches.
//
// Shape of the work: scores[b, r] = <q_b, c_r> (or -sum (q_b - c_r)^2 for l2) for every stored row,
// fp32 accumulation, and per query the k best (score desc, row asc).  No N-length score array is
// ever written: each warp keeps its running top-k in registers (WarpTopK), CTAs merge their warps
// in shared memory and write one k-list per (CTA, query); K3 (merge_topk.cu) merges those.
//
// Data movement (the roofline that bounds this kernel is HBM: 768 B/row bf16, 1536 B/row fp32):
//   * the corpus is read exactly once with 128-bit ld.global.nc.L1::no_allocate loads; a warp
//     reads 1536 contiguous bytes per "unit" (lane l takes 16-byte chunks l, l+32, l+64), four
//     units (6 KB, 12 loads per lane) are in flight per warp before any is consumed;
//   * the lane's slice of every query (3 chunks x 8 or 4 elements) stays in registers for the
//     whole kernel, so the inner loop is LDG + unpack + FFMA only;
//   * cross-lane reduction is a transposing butterfly: 32 rows x 32 lane-partials are reduced
//     with 31 shuffles per query (not 5 per row), leaving lane l with the finished score of row
//     bitrev5(l) of the 32-row block, so the top-k gate is one compare per lane.
//
// Fast path: rows of exactly 1536 B / 2 (bf16 x 384) or 1536 B (fp32 x 384, bf16 x 768).
// Any other dim (multiple of 8) takes scan_generic_kernel below.
#include "fr_common.cuh"
#include "fr_kernels.h"

namespace fr {

// ---------------------------------------------------------------------------------------------
// per-chunk partial dot products
template <bool L2>
__device__ __forceinline__ float dot_chunk_bf16(const uint4 &c, const float (&q)[8], float acc) {
    const uint32_t w[4] = {c.x, c.y, c.z, c.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float lo = __uint_as_float(w[i] << 16);
        const float hi = __uint_as_float(w[i] & 0xffff0000u);
        if (L2) {
            const float d0 = lo - q[2 * i], d1 = hi - q[2 * i + 1];
            acc = fmaf(d0, d0, acc);
            acc = fmaf(d1, d1, acc);
        } else {
            acc = fmaf(lo, q[2 * i], acc);
            acc = fmaf(hi, q[2 * i + 1], acc);
        }
    }
    return acc;
}

template <bool L2>
__device__ __forceinline__ float dot_chunk_f32(const uint4 &c, const float (&q)[8], float acc) {
    const float v[4] = {__uint_as_float(c.x), __uint_as_float(c.y), __uint_as_float(c.z),
                        __uint_as_float(c.w)};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        if (L2) {
            const float d = v[i] - q[i];
            acc = fmaf(d, d, acc);
        } else {
            acc = fmaf(v[i], q[i], acc);
        }
    }
    return acc;
}

__device__ __forceinline__ float bfly(float lo, float hi, int lane, int stride) {
    const bool up = (lane & stride) != 0;
    const float keep = up ? hi : lo;
    const float send = up ? lo : hi;
    return keep + __shfl_xor_sync(FULL_MASK, send, stride);
}

// ---------------------------------------------------------------------------------------------
// Fast kernel.  RU = rows per 1536-byte unit (2: bf16 x 384, 1: fp32 x 384 / bf16 x 768).
// BT = queries per pass (1, 2, 4), KPL = top-k slots per lane (k <= 32*KPL).
template <bool BF16, int RU, int BT, int KPL, bool L2>
__device__ __forceinline__ void
scan_stream_body(const uint8_t *__restrict__ corpus, const int64_t *__restrict__ keys_or_null,
                 const float *__restrict__ queries, int64_t n_rows, int k, int nq,
                 uint64_t *__restrict__ partials /* [gridDim.x][nq_total][k] */, int nq_total,
                 int q_offset) {
    constexpr int UNIT_BYTES = 1536;
    constexpr int ROW_BYTES = UNIT_BYTES / RU;
    constexpr int CPR = ROW_BYTES / 16;          // 16-byte chunks per row
    constexpr int EPC = BF16 ? 8 : 4;            // elements per chunk
    constexpr int DIM = CPR * EPC;
    constexpr int LEAVES_PER_GROUP = 4 * RU;     // 12 loads (4 x 1536 B) in flight per lane
    constexpr int UNITS_PER_GROUP = 4;
    constexpr int GROUPS = 32 / LEAVES_PER_GROUP;  // RU=2: 4 groups, RU=1: 8 groups

    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int nwarps = blockDim.x >> 5;

    // ---- this lane's slice of every query, kept in registers ----
    float q[BT][3][8];
#pragma unroll
    for (int b = 0; b < BT; ++b) {
        const int qb = (b < nq) ? b : (nq - 1);  // replicate the last query into unused slots
        const float *qp = queries + static_cast<size_t>(q_offset + qb) * DIM;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const int koff = ((lane + 32 * c) % CPR) * EPC;
#pragma unroll
            for (int e = 0; e < 8; ++e) q[b][c][e] = (e < EPC) ? qp[koff + e] : 0.0f;
        }
    }

    WarpTopK<KPL> tk[BT];
    float tau[BT];
#pragma unroll
    for (int b = 0; b < BT; ++b) {
        tk[b].clear();
        tau[b] = -INFINITY;
    }

    const int64_t n_blocks = (n_rows + 31) >> 5;
    const int64_t gwarp = static_cast<int64_t>(blockIdx.x) * nwarps + warp;
    const int64_t gstride = static_cast<int64_t>(gridDim.x) * nwarps;
    // bit-reversed lane = the row (leaf) of the 32-row block whose finished score lands here
    const uint32_t my_leaf = __brev(static_cast<uint32_t>(lane)) >> 27;

    for (int64_t blk = gwarp; blk < n_blocks; blk += gstride) {
        const int64_t row0 = blk << 5;
        const uint8_t *base = corpus + row0 * ROW_BYTES + lane * 16;
        float top[BT];     // result of the stride-1 butterfly
        float keep2[BT];   // pending value of the stride-2 level
        float keep1[BT];   // pending value of the stride-1 level
        float keep4[BT];   // RU==1 only: pending stride-4 level

#pragma unroll
        for (int g = 0; g < GROUPS; ++g) {
            // ---- issue all 12 loads of this group (6 KB contiguous per warp) ----
            uint4 ld[UNITS_PER_GROUP][3];
#pragma unroll
            for (int u = 0; u < UNITS_PER_GROUP; ++u)
#pragma unroll
                for (int c = 0; c < 3; ++c)
                    ld[u][c] = ld_stream_u4(base + (g * UNITS_PER_GROUP + u) * UNIT_BYTES + c * 512);

            // ---- per unit: lane-partial sums of its leaves, then butterfly up the tree ----
            float lvl16[UNITS_PER_GROUP][BT];  // RU=2: after stride 16 (one per unit)
#pragma unroll
            for (int u = 0; u < UNITS_PER_GROUP; ++u) {
#pragma unroll
                for (int b = 0; b < BT; ++b) {
                    float c0, c1, c2;
                    if (BF16) {
                        c0 = dot_chunk_bf16<L2>(ld[u][0], q[b][0], 0.0f);
                        c1 = dot_chunk_bf16<L2>(ld[u][1], q[b][1], 0.0f);
                        c2 = dot_chunk_bf16<L2>(ld[u][2], q[b][2], 0.0f);
                    } else {
                        c0 = dot_chunk_f32<L2>(ld[u][0], q[b][0], 0.0f);
                        c1 = dot_chunk_f32<L2>(ld[u][1], q[b][1], 0.0f);
                        c2 = dot_chunk_f32<L2>(ld[u][2], q[b][2], 0.0f);
                    }
                    if (RU == 2) {
                        // chunk 0 -> row 0; chunk 1 -> row 0 (lanes 0-15) / row 1 (16-31); chunk 2 -> row 1
                        const bool up = (lane & 16) != 0;
                        const float keep = c1 + (up ? c2 : c0);
                        const float send = up ? c0 : c2;
                        lvl16[u][b] = keep + __shfl_xor_sync(FULL_MASK, send, 16);
                    } else {
                        lvl16[u][b] = (c0 + c1) + c2;  // one leaf: all three chunks are the same row
                    }
                }
            }
#pragma unroll
            for (int b = 0; b < BT; ++b) {
                if (RU == 2) {
                    // 4 units = 8 leaves: strides 8, 4 inside the group
                    const float a = bfly(lvl16[0][b], lvl16[1][b], lane, 8);
                    const float c = bfly(lvl16[2][b], lvl16[3][b], lane, 8);
                    const float v4 = bfly(a, c, lane, 4);
                    // groups 0..3: stride 2 joins (0,1) and (2,3); stride 1 joins the halves
                    if ((g & 1) == 0) {
                        keep2[b] = v4;
                    } else {
                        const float v2 = bfly(keep2[b], v4, lane, 2);
                        if ((g & 2) == 0) keep1[b] = v2;
                        else top[b] = bfly(keep1[b], v2, lane, 1);
                    }
                } else {
                    // 4 units = 4 leaves: strides 16, 8 inside the group
                    const float a = bfly(lvl16[0][b], lvl16[1][b], lane, 16);
                    const float c = bfly(lvl16[2][b], lvl16[3][b], lane, 16);
                    const float v8 = bfly(a, c, lane, 8);
                    // groups 0..7: stride 4 joins pairs, stride 2 joins quads, stride 1 the halves
                    if ((g & 1) == 0) {
                        keep4[b] = v8;
                    } else {
                        const float v4 = bfly(keep4[b], v8, lane, 4);
                        if ((g & 2) == 0) {
                            keep2[b] = v4;
                        } else {
                            const float v2 = bfly(keep2[b], v4, lane, 2);
                            if ((g & 4) == 0) keep1[b] = v2;
                            else top[b] = bfly(keep1[b], v2, lane, 1);
                        }
                    }
                }
            }
        }

        // ---- fused top-k: lane holds the score of row row0 + my_leaf ----
        const int64_t my_row = row0 + my_leaf;
        const bool valid = my_row < n_rows;
#pragma unroll
        for (int b = 0; b < BT; ++b) {
            const float s = L2 ? -top[b] : top[b];
            unsigned m = __ballot_sync(FULL_MASK, valid && (s >= tau[b]));
            while (m) {
                const int src = __ffs(m) - 1;
                m &= m - 1;
                const float sv = __shfl_sync(FULL_MASK, s, src);
                const uint32_t rv = static_cast<uint32_t>(row0) + (__brev(static_cast<uint32_t>(src)) >> 27);
                if (keys_or_null != nullptr && keys_or_null[rv] == KEY_TOMBSTONE) continue;  // deleted row
                tk[b].insert(pack_key(sv, rv), k, lane);
                tau[b] = key_threshold(tk[b].kth(k));
            }
        }
    }

    // ---- CTA merge and one k-list per (CTA, query) ----
    extern __shared__ uint64_t smem_lists[];
    cta_merge_lists<KPL, BT>(tk, smem_lists, nwarps, warp, lane, k);
    if (warp == 0) {
#pragma unroll
        for (int b = 0; b < BT; ++b) {
            if (b < nq) {
                uint64_t *dst = partials + (static_cast<size_t>(blockIdx.x) * nq_total + q_offset + b) * k;
#pragma unroll
                for (int j = 0; j < KPL; ++j) {
                    const int i = j * 32 + lane;
                    if (i < k) dst[i] = tk[b].e[j];
                }
            }
        }
    }
}

template <bool BF16, int RU, int BT, int KPL, bool L2>
__global__ void __launch_bounds__(256)
scan_stream_kernel(const uint8_t *__restrict__ corpus, const int64_t *__restrict__ keys_or_null,
                   const float *__restrict__ queries, int64_t n_rows, int k, int nq,
                   uint64_t *__restrict__ partials, int nq_total, int q_offset) {
    scan_stream_body<BF16, RU, BT, KPL, L2>(corpus, keys_or_null, queries, n_rows, k, nq, partials, nq_total, q_offset);
}

// K2's safety net: re-scan, one query per corpus pass, the queries whose tensor-core selection could
// not be certified (scan_mma.cu).  Launched unconditionally; returns at once when nothing failed.
template <bool BF16, int RU, int KPL, bool L2>
__global__ void __launch_bounds__(256)
scan_stream_fallback_kernel(const uint8_t *__restrict__ corpus, const int64_t *__restrict__ keys_or_null,
                            const float *__restrict__ queries, int64_t n_rows, int k,
                            uint64_t *__restrict__ partials, int nq_total, const int *__restrict__ fail_count,
                            const int *__restrict__ fail_list) {
    const int nf = *fail_count;
    for (int f = 0; f < nf; ++f) {
        scan_stream_body<BF16, RU, 1, KPL, L2>(corpus, keys_or_null, queries, n_rows, k, 1, partials, nq_total,
                                               fail_list[f]);
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------------
// Generic kernel: any dim that is a multiple of 8 (bf16) / 4 (fp32).  One warp per row, lanes
// stride over 16-byte chunks, queries staged in shared memory.  Correctness path for the
// multi-vector store's other widths (multivector_store.py:70); not the tuned one.
template <bool BF16, int BT, int KPL, bool L2>
__global__ void __launch_bounds__(256)
scan_generic_kernel(const uint8_t *__restrict__ corpus, const int64_t *__restrict__ keys_or_null,
                    const float *__restrict__ queries, int64_t n_rows, int dim, int k, int nq,
                    uint64_t *__restrict__ partials, int nq_total, int q_offset) {
    extern __shared__ uint64_t smem_lists[];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int nwarps = blockDim.x >> 5;
    constexpr int EPC = BF16 ? 8 : 4;
    const int cpr = dim / EPC;
    const size_t row_bytes = static_cast<size_t>(dim) * (BF16 ? 2 : 4);
    // queries live after the merge lists in dynamic shared memory
    float *sq = reinterpret_cast<float *>(smem_lists + static_cast<size_t>(nwarps) * BT * 32 * KPL);
    for (int i = threadIdx.x; i < BT * dim; i += blockDim.x) {
        const int b = i / dim, e = i - b * dim;
        const int qb = (b < nq) ? b : (nq - 1);
        sq[i] = queries[static_cast<size_t>(q_offset + qb) * dim + e];
    }
    __syncthreads();

    WarpTopK<KPL> tk[BT];
    float tau[BT];
#pragma unroll
    for (int b = 0; b < BT; ++b) {
        tk[b].clear();
        tau[b] = -INFINITY;
    }
    const int64_t gwarp = static_cast<int64_t>(blockIdx.x) * nwarps + warp;
    const int64_t gstride = static_cast<int64_t>(gridDim.x) * nwarps;
    for (int64_t row = gwarp; row < n_rows; row += gstride) {
        const uint8_t *rp = corpus + row * row_bytes;
        float acc[BT];
#pragma unroll
        for (int b = 0; b < BT; ++b) acc[b] = 0.0f;
        for (int c = lane; c < cpr; c += 32) {
            const uint4 v = ld_stream_u4(rp + static_cast<size_t>(c) * 16);
#pragma unroll
            for (int b = 0; b < BT; ++b) {
                float qq[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) qq[e] = (e < EPC) ? sq[b * dim + c * EPC + e] : 0.0f;
                acc[b] = BF16 ? dot_chunk_bf16<L2>(v, qq, acc[b]) : dot_chunk_f32<L2>(v, qq, acc[b]);
            }
        }
#pragma unroll
        for (int b = 0; b < BT; ++b) {
#pragma unroll
            for (int s = 16; s > 0; s >>= 1) acc[b] += __shfl_xor_sync(FULL_MASK, acc[b], s);
            const float sc = L2 ? -acc[b] : acc[b];
            if (sc >= tau[b]) {  // warp-uniform
                if (keys_or_null == nullptr || keys_or_null[row] != KEY_TOMBSTONE) {
                    tk[b].insert(pack_key(sc, static_cast<uint32_t>(row)), k, lane);
                    tau[b] = key_threshold(tk[b].kth(k));
                }
            }
        }
    }
    cta_merge_lists<KPL, BT>(tk, smem_lists, nwarps, warp, lane, k);
    if (warp == 0) {
#pragma unroll
        for (int b = 0; b < BT; ++b) {
            if (b < nq) {
                uint64_t *dst = partials + (static_cast<size_t>(blockIdx.x) * nq_total + q_offset + b) * k;
#pragma unroll
                for (int j = 0; j < KPL; ++j) {
                    const int i = j * 32 + lane;
                    if (i < k) dst[i] = tk[b].e[j];
                }
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// host-side dispatch.  Every launcher doubles as an occupancy probe (occ_out != nullptr): the
// persistent grid must not exceed the resident-CTA count or the grid-stride loop runs in waves.
namespace {

constexpr int THREADS = 256;

template <typename K>
cudaError_t prepare_kernel(K kernel, size_t smem, int *occ_out) {
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             static_cast<int>(smem));
        if (e != cudaSuccess) return e;
    }
    if (occ_out) {
        int per_sm = 0;
        cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, THREADS, smem);
        if (e != cudaSuccess) return e;
        if (per_sm < 1) per_sm = 1;
        if (per_sm < *occ_out) *occ_out = per_sm;
    }
    return cudaSuccess;
}

template <bool BF16, int RU, int BT, int KPL, bool L2>
cudaError_t launch_fast(const ScanArgs &a, int q_offset, int nq, int *occ_out) {
    const size_t smem = static_cast<size_t>(THREADS / 32) * BT * 32 * KPL * sizeof(uint64_t);
    auto kern = scan_stream_kernel<BF16, RU, BT, KPL, L2>;
    cudaError_t e = prepare_kernel(kern, smem, occ_out);
    if (e != cudaSuccess || occ_out) return e;
    kern<<<a.grid, THREADS, smem, a.stream>>>(a.corpus, a.keys_or_null, a.queries, a.n_rows, a.k, nq,
                                              a.partials, a.nq_total, q_offset);
    count_launch();
    return cudaGetLastError();
}

template <bool BF16, int BT, int KPL, bool L2>
cudaError_t launch_generic(const ScanArgs &a, int q_offset, int nq, int *occ_out) {
    const size_t smem = static_cast<size_t>(THREADS / 32) * BT * 32 * KPL * sizeof(uint64_t) +
                        static_cast<size_t>(BT) * a.dim * sizeof(float);
    auto kern = scan_generic_kernel<BF16, BT, KPL, L2>;
    cudaError_t e = prepare_kernel(kern, smem, occ_out);
    if (e != cudaSuccess || occ_out) return e;
    kern<<<a.grid, THREADS, smem, a.stream>>>(a.corpus, a.keys_or_null, a.queries, a.n_rows, a.dim, a.k, nq,
                                              a.partials, a.nq_total, q_offset);
    count_launch();
    return cudaGetLastError();
}

template <bool BF16, int BT, int KPL>
cudaError_t launch_bt(const ScanArgs &a, int q_offset, int nq, int *occ_out) {
    const int row_bytes = a.dim * (BF16 ? 2 : 4);
    const bool l2 = a.l2;
    if (BF16 && row_bytes == 768) {  // bf16 x 384: two rows per 1536-byte unit
        return l2 ? launch_fast<BF16, 2, BT, KPL, true>(a, q_offset, nq, occ_out)
                  : launch_fast<BF16, 2, BT, KPL, false>(a, q_offset, nq, occ_out);
    }
    if (row_bytes == 1536) {  // fp32 x 384 or bf16 x 768: one row per unit
        return l2 ? launch_fast<BF16, 1, BT, KPL, true>(a, q_offset, nq, occ_out)
                  : launch_fast<BF16, 1, BT, KPL, false>(a, q_offset, nq, occ_out);
    }
    return l2 ? launch_generic<BF16, BT, KPL, true>(a, q_offset, nq, occ_out)
              : launch_generic<BF16, BT, KPL, false>(a, q_offset, nq, occ_out);
}

// query tiles of 4, then 2, then 1 (each tile is one pass over the corpus)
template <bool BF16, int KPL>
cudaError_t launch_tiles(const ScanArgs &a, int *occ_out) {
    int off = 0;
    bool seen4 = false;
    while (off < a.nq_total) {
        const int left = a.nq_total - off;
        cudaError_t e = cudaSuccess;
        if (left >= 4) {
            if (!(occ_out && seen4)) e = launch_bt<BF16, 4, KPL>(a, off, 4, occ_out);
            seen4 = true;
            off += 4;
        } else if (left >= 2) {
            e = launch_bt<BF16, 2, KPL>(a, off, 2, occ_out);
            off += 2;
        } else {
            e = launch_bt<BF16, 1, KPL>(a, off, 1, occ_out);
            off += 1;
        }
        if (e != cudaSuccess) return e;
    }
    return cudaSuccess;
}

cudaError_t dispatch(const ScanArgs &a, int *occ_out) {
    if (a.k <= 32) return a.bf16 ? launch_tiles<true, 1>(a, occ_out) : launch_tiles<false, 1>(a, occ_out);
    return a.bf16 ? launch_tiles<true, 4>(a, occ_out) : launch_tiles<false, 4>(a, occ_out);
}

}  // namespace

int scan_stream_plan_grid(const ScanArgs &a, int sm_count) {
    int occ = 1 << 20;
    if (dispatch(a, &occ) != cudaSuccess || occ == (1 << 20)) occ = 1;
    // one warp handles 32 rows per iteration; do not launch CTAs that would have no block
    const int row_bytes = a.dim * (a.bf16 ? 2 : 4);
    const bool fast = (a.bf16 && row_bytes == 768) || row_bytes == 1536;
    const int64_t work_items = fast ? (a.n_rows + 31) / 32 : a.n_rows;
    int64_t want = (work_items + (THREADS / 32) - 1) / (THREADS / 32);
    if (want < 1) want = 1;
    const int64_t cap = static_cast<int64_t>(sm_count) * occ;
    return static_cast<int>(want < cap ? want : cap);
}

cudaError_t launch_scan_stream(const ScanArgs &a) { return dispatch(a, nullptr); }

// fallback: bf16 rows of 768 B (width 384, two per unit) or 1536 B (width 768), or fp32 rows of 1536 B (width 384, the
// collections whose tensor-core scans read a bf16 copy); dot metrics only
namespace {
template <int KPL, int RU, bool BF16, bool L2 = false>
cudaError_t fallback_impl(const ScanArgs &a, const int *fail_count, const int *fail_list, int *occ_out) {
    const size_t smem = static_cast<size_t>(THREADS / 32) * 32 * KPL * sizeof(uint64_t);
    auto kern = scan_stream_fallback_kernel<BF16, RU, KPL, L2>;
    cudaError_t e = prepare_kernel(kern, smem, occ_out);
    if (e != cudaSuccess || occ_out) return e;
    kern<<<a.grid, THREADS, smem, a.stream>>>(a.corpus, a.keys_or_null, a.queries, a.n_rows, a.k, a.partials,
                                              a.nq_total, fail_count, fail_list);
    count_launch();
    return cudaGetLastError();
}
cudaError_t fallback_dispatch(const ScanArgs &a, const int *fail_count, const int *fail_list, int *occ_out) {
    if (!a.bf16)  // fp32 x 384: one row per 1536-byte unit
        return a.k <= 32 ? fallback_impl<1, 1, false>(a, fail_count, fail_list, occ_out)
                         : fallback_impl<4, 1, false>(a, fail_count, fail_list, occ_out);
    if (a.dim == 384 && a.l2)  // l2 collections of width 384 (the K2s selection's safety net)
        return a.k <= 32 ? fallback_impl<1, 2, true, true>(a, fail_count, fail_list, occ_out)
                         : fallback_impl<4, 2, true, true>(a, fail_count, fail_list, occ_out);
    if (a.dim == 384)
        return a.k <= 32 ? fallback_impl<1, 2, true>(a, fail_count, fail_list, occ_out)
                         : fallback_impl<4, 2, true>(a, fail_count, fail_list, occ_out);
    return a.k <= 32 ? fallback_impl<1, 1, true>(a, fail_count, fail_list, occ_out)
                     : fallback_impl<4, 1, true>(a, fail_count, fail_list, occ_out);
}
}  // namespace

bool scan_stream_fallback_serves(const ScanArgs &a) {
    if (a.l2) return a.bf16 && a.dim == 384;
    return a.bf16 ? (a.dim == 384 || a.dim == 768) : a.dim == 384;
}

int scan_stream_fallback_grid(const ScanArgs &a, int sm_count) {
    int occ = 1 << 20;
    cudaError_t e = fallback_dispatch(a, nullptr, nullptr, &occ);
    if (e != cudaSuccess || occ == (1 << 20)) occ = 1;
    int64_t want = ((a.n_rows + 31) / 32 + (THREADS / 32) - 1) / (THREADS / 32);
    if (want < 1) want = 1;
    const int64_t cap = static_cast<int64_t>(sm_count) * occ;
    return static_cast<int>(want < cap ? want : cap);
}

cudaError_t launch_scan_stream_fallback(const ScanArgs &a, const int *fail_count, const int *fail_list) {
    return fallback_dispatch(a, fail_count, fail_list, nullptr);
}

}  // namespace fr
