// Query-side encoder on the GPU: the BERT forward pass of the reference's two embedders --
// BAAI/bge-small-en-v1.5 (CLS pooling) and thenlper/gte-small (mean pooling), both 12-layer BERT-384
// (local_models/*/config.json, local_models/*/1_Pooling/config.json) -- so that a query goes from token ids to the
// normalised 384-d vector the scan reads without leaving the device.
// Replaces what SentenceTransformer.encode / local_embedder.py:155-191 do per query on the CPU:
//     embeddings (word + position + token type, LayerNorm) -> 12 x [self-attention, add & LayerNorm, GELU MLP, add &
//     LayerNorm] -> pooling (CLS | masked mean, local_embedder.py:171-179) -> L2 normalisation (:182).
//
// Kernels (sm_100a):
//   E1 embed_ln_kernel      warp per token: three table rows summed, LayerNorm (fp32)
//   E2 encoder_gemm_kernel  Y[M x N] = X[M x K] W[N x K]^T + b on the 5th-gen tensor cores: TMA -> shared-memory ring ->
//                           tcgen05.mma (bf16, fp32 accumulation in TMEM) -> tcgen05.ld epilogue (bias, GELU).
//                           fp32 fidelity from bf16 tensor cores: every fp32 operand is kept as TWO bf16 terms
//                           (x = hi + lo, lo = bf16(x - hi): 16 mantissa bits) and the product is accumulated as
//                           hi*hi + hi*lo + lo*hi -- three passes over K into the same accumulator; the dropped lo*lo
//                           term is ~2^-16 relative.  The model is tiny (21M layer parameters): the tripled tensor work
//                           is noise, and the embeddings stay within 1e-4 of an fp32 reference (tests: 1e-3).
//   E3 attention_kernel     CTA per (sequence, head), fp32: K/V of the sequence in shared memory, a warp per query row
//                           (scores, softmax by warp reductions, P V), sequences up to 512 tokens
//   E4 add_ln_kernel        warp per token: residual add + LayerNorm, writes fp32 and the two bf16 terms the next GEMM reads
//   E5 pool_kernel          CTA per sequence: CLS or masked mean, L2 normalisation
// Roofline: latency.  A query batch is a few hundred tokens; one forward pass is 86 short launches.
#include <cstring>
#include <mutex>
#include <new>
#include <string>
#include <unordered_map>
#include <vector>

#include "fr_host.h"
#include "fr_kernels.h"
#include "mma_common.cuh"

namespace fr {
namespace enc {

using namespace fr::mma;

constexpr int G_M = 128;        // tokens per CTA tile (one TMEM lane each)
constexpr int G_N = 128;        // output features per CTA tile (TMEM columns)
constexpr int G_STAGES = 4;
constexpr int G_A_BYTES = G_M * K_CHUNK * 2;            // 16 KB
constexpr int G_B_BYTES = G_N * K_CHUNK * 2;            // 16 KB
constexpr int G_STAGE_BYTES = G_A_BYTES + G_B_BYTES;
constexpr size_t G_SMEM = size_t(G_STAGES) * G_STAGE_BYTES + 256 + 1024;
enum { EPI_F32 = 0, EPI_GELU_SPLIT = 1 };

__device__ __forceinline__ void split_bf16(float x, __nv_bfloat16 &hi, __nv_bfloat16 &lo) {
    hi = __float2bfloat16_rn(x);
    lo = __float2bfloat16_rn(x - __bfloat162float(hi));
}

// fp32 -> the two bf16 terms (weights, once at load time)
__global__ void split_kernel(const float *__restrict__ src, __nv_bfloat16 *__restrict__ hi, __nv_bfloat16 *__restrict__ lo,
                             int64_t n) {
    for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n;
         i += static_cast<int64_t>(gridDim.x) * blockDim.x)
        split_bf16(src[i], hi[i], lo[i]);
}

// ---- E2: Y = X W^T + b ----------------------------------------------------------------------------------------------
// grid (ceil(M / 128), N / 128); 192 threads: warp 0 = TMA producer, warp 1 = MMA issuer + TMEM owner, warps 2-5 = epilogue
// (warp w reads TMEM lanes [32 (w % 4), +32): a thread per token).  K runs over three passes of k_chunks 64-element chunks:
// (X_hi, W_hi), (X_hi, W_lo), (X_lo, W_hi).
template <int EPI>
__global__ void __launch_bounds__(192, 1)
encoder_gemm_kernel(const __grid_constant__ CUtensorMap tm_xh, const __grid_constant__ CUtensorMap tm_xl,
                    const __grid_constant__ CUtensorMap tm_wh, const __grid_constant__ CUtensorMap tm_wl, int M, int k_chunks,
                    const float *__restrict__ bias, float *__restrict__ out_f32, __nv_bfloat16 *__restrict__ out_hi,
                    __nv_bfloat16 *__restrict__ out_lo, int out_pitch) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + size_t(G_STAGES) * G_STAGE_BYTES);
    const uint32_t bar_full = smem_u32(bars), bar_empty = smem_u32(bars + G_STAGES), bar_done = smem_u32(bars + 2 * G_STAGES);
    uint32_t *tmem_ptr_smem = reinterpret_cast<uint32_t *>(bars + 2 * G_STAGES + 1);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m0 = blockIdx.x * G_M, n0 = blockIdx.y * G_N;
    const int total = 3 * k_chunks;

    if (threadIdx.x == 0) {
        for (int s = 0; s < G_STAGES; ++s) {
            mbar_init(bar_full + 8 * s, 1);
            mbar_init(bar_empty + 8 * s, 1);
        }
        mbar_init(bar_done, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr_smem)), "r"(G_N)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;

    if (warp == 0) {
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int it = 0; it < total; ++it) {
                const int term = it / k_chunks, kc = it - term * k_chunks;
                const CUtensorMap *mx = term < 2 ? &tm_xh : &tm_xl;
                const CUtensorMap *mw = term == 1 ? &tm_wl : &tm_wh;
                mbar_wait(bar_empty + 8 * stage, phase ^ 1);
                mbar_expect_tx(bar_full + 8 * stage, G_STAGE_BYTES);
                uint8_t *dst = smem + size_t(stage) * G_STAGE_BYTES;
                tma_load_2d<1>(smem_u32(dst), mx, bar_full + 8 * stage, kc * K_CHUNK, m0);
                tma_load_2d<1>(smem_u32(dst + G_A_BYTES), mw, bar_full + 8 * stage, kc * K_CHUNK, n0);
                if (++stage == G_STAGES) {
                    stage = 0;
                    phase ^= 1;
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            constexpr uint32_t idesc = make_idesc_bf16(G_M, G_N);
            int stage = 0;
            uint32_t phase = 0;
            for (int it = 0; it < total; ++it) {
                mbar_wait(bar_full + 8 * stage, phase);
                tc_fence_after();
                const uint32_t a_addr = smem_u32(smem + size_t(stage) * G_STAGE_BYTES);
                const uint32_t b_addr = a_addr + G_A_BYTES;
#pragma unroll
                for (int k4 = 0; k4 < K_CHUNK / UMMA_K; ++k4)
                    tc_mma_bf16<1>(tmem_base, make_kmajor_sw128_desc(a_addr + k4 * UMMA_K * 2),
                                   make_kmajor_sw128_desc(b_addr + k4 * UMMA_K * 2), idesc, (it | k4) != 0 ? 1u : 0u);
                tc_commit<1>(bar_empty + 8 * stage);
                if (++stage == G_STAGES) {
                    stage = 0;
                    phase ^= 1;
                }
            }
            tc_commit<1>(bar_done);
        }
    } else {
        const int quarter = warp & 3;
        const int row = m0 + quarter * 32 + lane;
        mbar_wait(bar_done, 0);
        tc_fence_after();
#pragma unroll 1
        for (int c0 = 0; c0 < G_N; c0 += 64) {
            float v[64];
            tmem_ld64(tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + c0, v);  // warp-collective: every lane
            const float *bp = bias + n0 + c0;
            if (row >= M) {
                // a token past the end of the batch: nothing to store (the loads above are collective, so it took part)
            } else if constexpr (EPI == EPI_F32) {
                float4 *dst = reinterpret_cast<float4 *>(out_f32 + static_cast<size_t>(row) * out_pitch + n0 + c0);
#pragma unroll
                for (int c = 0; c < 64; c += 4)
                    dst[c / 4] = make_float4(v[c] + bp[c], v[c + 1] + bp[c + 1], v[c + 2] + bp[c + 2], v[c + 3] + bp[c + 3]);
            } else {
                uint4 *dh = reinterpret_cast<uint4 *>(out_hi + static_cast<size_t>(row) * out_pitch + n0 + c0);
                uint4 *dl = reinterpret_cast<uint4 *>(out_lo + static_cast<size_t>(row) * out_pitch + n0 + c0);
#pragma unroll
                for (int c = 0; c < 64; c += 8) {
                    __nv_bfloat16 h[8], l[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const float y = v[c + i] + bp[c + i];
                        const float g = 0.5f * y * (1.0f + erff(y * 0.70710678118654752f));  // BERT's "gelu" (erf form)
                        split_bf16(g, h[i], l[i]);
                    }
                    dh[c / 8] = *reinterpret_cast<const uint4 *>(h);
                    dl[c / 8] = *reinterpret_cast<const uint4 *>(l);
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(G_N) : "memory");
    }
}

// ---- LayerNorm of one token held by a warp (x[] = the lane's elements e = lane, lane + 32, ...) ----------------------
template <int MAXE>
__device__ __forceinline__ void warp_layernorm_store(float (&x)[MAXE], int H, int lane, float eps, const float *__restrict__ gamma,
                                                     const float *__restrict__ beta, float *__restrict__ of32,
                                                     __nv_bfloat16 *__restrict__ ohi, __nv_bfloat16 *__restrict__ olo) {
    float s = 0.0f;
#pragma unroll
    for (int i = 0; i < MAXE; ++i)
        if (lane + 32 * i < H) s += x[i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(FULL_MASK, s, o);
    const float mean = s / static_cast<float>(H);
    float q = 0.0f;
#pragma unroll
    for (int i = 0; i < MAXE; ++i)
        if (lane + 32 * i < H) {
            const float d = x[i] - mean;
            q += d * d;
        }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(FULL_MASK, q, o);
    const float rstd = rsqrtf(q / static_cast<float>(H) + eps);
#pragma unroll
    for (int i = 0; i < MAXE; ++i) {
        const int e = lane + 32 * i;
        if (e < H) {
            const float y = (x[i] - mean) * rstd * gamma[e] + beta[e];
            of32[e] = y;
            split_bf16(y, ohi[e], olo[e]);
        }
    }
}

constexpr int LN_MAXE = 32;  // hidden <= 1024

// ---- E1: embeddings + LayerNorm.  grid = ceil(M / 4), 128 threads (a warp per token) ----------------------------------
__global__ void __launch_bounds__(128)
embed_ln_kernel(const int32_t *__restrict__ ids, int M, int T, int H, int vocab, const float *__restrict__ word,
                const float *__restrict__ pos, const float *__restrict__ type0, const float *__restrict__ gamma,
                const float *__restrict__ beta, float eps, float *__restrict__ of32, __nv_bfloat16 *__restrict__ ohi,
                __nv_bfloat16 *__restrict__ olo) {
    const int tok = blockIdx.x * 4 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (tok >= M) return;
    int id = ids[tok];
    id = id < 0 ? 0 : (id >= vocab ? vocab - 1 : id);
    const int p = tok % T;
    float x[LN_MAXE];
#pragma unroll
    for (int i = 0; i < LN_MAXE; ++i) {
        const int e = lane + 32 * i;
        x[i] = e < H ? word[static_cast<size_t>(id) * H + e] + type0[e] + pos[static_cast<size_t>(p) * H + e] : 0.0f;
    }
    const size_t o = static_cast<size_t>(tok) * H;
    warp_layernorm_store<LN_MAXE>(x, H, lane, eps, gamma, beta, of32 + o, ohi + o, olo + o);
}

// ---- E4: hidden = LayerNorm(y + hidden) -----------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
add_ln_kernel(const float *__restrict__ y, int M, int H, const float *__restrict__ gamma, const float *__restrict__ beta,
              float eps, float *__restrict__ hid, __nv_bfloat16 *__restrict__ ohi, __nv_bfloat16 *__restrict__ olo) {
    const int tok = blockIdx.x * 4 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (tok >= M) return;
    const size_t o = static_cast<size_t>(tok) * H;
    float x[LN_MAXE];
#pragma unroll
    for (int i = 0; i < LN_MAXE; ++i) {
        const int e = lane + 32 * i;
        x[i] = e < H ? y[o + e] + hid[o + e] : 0.0f;
    }
    warp_layernorm_store<LN_MAXE>(x, H, lane, eps, gamma, beta, hid + o, ohi + o, olo + o);
}

// ---- E3: self-attention of one (sequence, head) ------------------------------------------------------------------------
// qkv: [M][3H] fp32 (Q | K | V, head h at columns h * HD).  Keys j >= len are masked out (right padding); query rows >= len
// produce values nobody reads (pooling looks at valid tokens only).  Shared memory: K and V of the sequence, rows padded to
// HD + 1 floats (a lane reads row `lane`: without the pad all 32 lanes hit one bank), plus a probability row per warp.
template <int HD>
__global__ void __launch_bounds__(128)
attention_kernel(const float *__restrict__ qkv, const int32_t *__restrict__ lens, int T, int H, float scale,
                 __nv_bfloat16 *__restrict__ ctx_hi, __nv_bfloat16 *__restrict__ ctx_lo) {
    extern __shared__ float sm[];
    const int b = blockIdx.x, h = blockIdx.y;
    const int len = min(max(lens[b], 1), T);
    float *sk = sm, *sv = sm + static_cast<size_t>(T) * (HD + 1), *sp = sv + static_cast<size_t>(T) * (HD + 1);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    const size_t pitch = static_cast<size_t>(3) * H;
    const float *base = qkv + static_cast<size_t>(b) * T * pitch + static_cast<size_t>(h) * HD;
    for (int i = threadIdx.x; i < len * HD; i += blockDim.x) {
        const int j = i / HD, d = i - j * HD;
        sk[j * (HD + 1) + d] = base[j * pitch + H + d];
        sv[j * (HD + 1) + d] = base[j * pitch + 2 * H + d];
    }
    __syncthreads();
    float *prob = sp + static_cast<size_t>(warp) * T;
    for (int i = warp; i < T; i += nw) {
        float q[HD];
#pragma unroll
        for (int d = 0; d < HD; ++d) q[d] = base[i * pitch + d] * scale;  // same address for the whole warp: broadcast
        float mx = -INFINITY;
        for (int j = lane; j < len; j += 32) {
            const float *kr = sk + j * (HD + 1);
            float s = 0.0f;
#pragma unroll
            for (int d = 0; d < HD; ++d) s = fmaf(q[d], kr[d], s);
            prob[j] = s;
            mx = fmaxf(mx, s);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(FULL_MASK, mx, o));
        float sum = 0.0f;
        for (int j = lane; j < len; j += 32) {
            const float e = expf(prob[j] - mx);
            prob[j] = e;
            sum += e;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(FULL_MASK, sum, o);
        __syncwarp();
        const float inv = 1.0f / sum;
        // lane d (and d + 32 for 64-wide heads) accumulates output dimension d over all keys
        float acc[(HD + 31) / 32];
#pragma unroll
        for (int r = 0; r < (HD + 31) / 32; ++r) acc[r] = 0.0f;
        for (int j = 0; j < len; ++j) {
            const float pj = prob[j];
#pragma unroll
            for (int r = 0; r < (HD + 31) / 32; ++r) {
                const int d = lane + 32 * r;
                if (d < HD) acc[r] = fmaf(pj, sv[j * (HD + 1) + d], acc[r]);
            }
        }
        const size_t o = (static_cast<size_t>(b) * T + i) * H + static_cast<size_t>(h) * HD;
#pragma unroll
        for (int r = 0; r < (HD + 31) / 32; ++r) {
            const int d = lane + 32 * r;
            if (d < HD) split_bf16(acc[r] * inv, ctx_hi[o + d], ctx_lo[o + d]);
        }
        __syncwarp();
    }
}

// ---- E5: pooling + L2 normalisation.  grid = B, 128 threads --------------------------------------------------------
// pooling 0: the [CLS] token (bge-small: 1_Pooling/config.json pooling_mode_cls_token); 1: mean over the valid tokens
// (gte-small: pooling_mode_mean_tokens; local_embedder.py:171-179).  normalize: x / max(|x|, 1e-12) (local_embedder.py:182).
__global__ void __launch_bounds__(128)
pool_kernel(const float *__restrict__ hid, const int32_t *__restrict__ lens, int T, int H, int pooling, int normalize,
            float *__restrict__ out) {
    __shared__ float red[4];
    const int b = blockIdx.x;
    const int len = min(max(lens[b], 1), T);
    float v[8];  // H <= 1024
    float ss = 0.0f;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int e = threadIdx.x + 128 * i;
        float x = 0.0f;
        if (e < H) {
            const float *p = hid + static_cast<size_t>(b) * T * H + e;
            if (pooling == 0) {
                x = p[0];
            } else {
                for (int t = 0; t < len; ++t) x += p[static_cast<size_t>(t) * H];
                x = x / static_cast<float>(len);
            }
        }
        v[i] = x;
        ss += x * x;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(FULL_MASK, ss, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = ss;
    __syncthreads();
    const float norm = sqrtf(red[0] + red[1] + red[2] + red[3]);
    const float inv = normalize ? 1.0f / fmaxf(norm, 1e-12f) : 1.0f;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int e = threadIdx.x + 128 * i;
        if (e < H) out[static_cast<size_t>(b) * H + e] = v[i] * inv;
    }
}

}  // namespace enc
}  // namespace fr

// =====================================================================================================================
namespace {

using fr::DevBuf;
using fr::DeviceGuard;
using fr::fail;
using fr::PinBuf;
typedef __nv_bfloat16 bf16;

struct SplitW {  // a Linear weight [N][K] as two bf16 terms + its fp32 bias
    DevBuf hi, lo, bias;
    int N = 0, K = 0;
    bool have_w = false, have_b = false;
    CUtensorMap map_hi, map_lo;
};
struct LnP {
    DevBuf gamma, beta;
    bool have_g = false, have_b = false;
};
struct Layer {
    SplitW qkv, out, ffn1, ffn2;
    int qkv_parts = 0, qkv_bias_parts = 0;  // bit mask of q / k / v loaded
    LnP ln1, ln2;
};

}  // namespace

struct fr_encoder {
    int device = 0, vocab = 0, H = 0, L = 0, heads = 0, I = 0, max_pos = 0, type_vocab = 0;
    float eps = 1e-12f;
    DevBuf word, pos, type;
    bool have_word = false, have_pos = false, have_type = false;
    LnP emb_ln;
    std::vector<Layer> layers;
    bool ready = false;
    // activations (grow-only, M_cap tokens)
    int64_t m_cap = 0;
    DevBuf hid, hid_hi, hid_lo, qkv, ctx_hi, ctx_lo, tmp, ffn_hi, ffn_lo, ids, lens, out, stage;
    CUtensorMap map_hid_hi, map_hid_lo, map_ctx_hi, map_ctx_lo, map_ffn_hi, map_ffn_lo;
    PinBuf pin;
    cudaStream_t stream = nullptr;
    std::mutex mu;
    // host entry point: one captured CUDA graph per (B, T, pooling, normalize) -- a forward pass is 86 short launches and
    // a query batch is latency-bound -- replayed while the buffers it baked in have not moved
    struct Graph {
        cudaGraphExec_t exec = nullptr;
        uint64_t state = 0, seen = 0;
        int64_t launches = 0;
        bool failed = false;
    };
    std::unordered_map<uint64_t, Graph> graphs;
    int64_t n_graph_replays = 0;
};

namespace {

int upload(DevBuf &dst, const float *src, int64_t n, cudaStream_t s) {
    FR_CUDA(dst.need(static_cast<size_t>(n) * sizeof(float)));
    FR_CUDA(cudaMemcpyAsync(dst.p, src, static_cast<size_t>(n) * sizeof(float), cudaMemcpyHostToDevice, s));
    FR_CUDA(cudaStreamSynchronize(s));
    return FR_OK;
}

// rows [row0, row0 + rows) of a [N][K] weight from host fp32 -> the two bf16 terms on the device
int upload_split_rows(fr_encoder *e, SplitW &w, int N, int K, int row0, int rows, const float *src) {
    if (w.N == 0) {
        w.N = N;
        w.K = K;
        FR_CUDA(w.hi.need(static_cast<size_t>(N) * K * 2));
        FR_CUDA(w.lo.need(static_cast<size_t>(N) * K * 2));
    }
    const int64_t n = static_cast<int64_t>(rows) * K;
    FR_CUDA(e->stage.need(static_cast<size_t>(n) * sizeof(float)));
    FR_CUDA(cudaMemcpyAsync(e->stage.p, src, static_cast<size_t>(n) * sizeof(float), cudaMemcpyHostToDevice, e->stream));
    fr::enc::split_kernel<<<256, 256, 0, e->stream>>>(static_cast<const float *>(e->stage.p),
                                                      static_cast<bf16 *>(w.hi.p) + static_cast<size_t>(row0) * K,
                                                      static_cast<bf16 *>(w.lo.p) + static_cast<size_t>(row0) * K, n);
    fr::count_launch();
    FR_CUDA(cudaGetLastError());
    FR_CUDA(cudaStreamSynchronize(e->stream));
    return FR_OK;
}

int upload_bias_rows(fr_encoder *e, SplitW &w, int N, int row0, int rows, const float *src) {
    FR_CUDA(w.bias.need(static_cast<size_t>(N) * sizeof(float)));
    FR_CUDA(cudaMemcpyAsync(static_cast<float *>(w.bias.p) + row0, src, static_cast<size_t>(rows) * sizeof(float),
                            cudaMemcpyHostToDevice, e->stream));
    FR_CUDA(cudaStreamSynchronize(e->stream));
    return FR_OK;
}

bool make_maps(SplitW &w) {
    return fr::mma::make_row_major_map(&w.map_hi, w.hi.p, w.N, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, fr::enc::G_N, w.K) &&
           fr::mma::make_row_major_map(&w.map_lo, w.lo.p, w.N, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, fr::enc::G_N, w.K);
}

int ensure_activations(fr_encoder *e, int64_t M) {
    if (M <= e->m_cap) return FR_OK;
    const int64_t cap = ((M + 127) / 128) * 128;
    const size_t H = e->H, I = e->I;
    FR_CUDA(cudaStreamSynchronize(e->stream));
    FR_CUDA(e->hid.need(cap * H * 4));
    FR_CUDA(e->hid_hi.need(cap * H * 2));
    FR_CUDA(e->hid_lo.need(cap * H * 2));
    FR_CUDA(e->qkv.need(cap * 3 * H * 4));
    FR_CUDA(e->ctx_hi.need(cap * H * 2));
    FR_CUDA(e->ctx_lo.need(cap * H * 2));
    FR_CUDA(e->tmp.need(cap * H * 4));
    FR_CUDA(e->ffn_hi.need(cap * I * 2));
    FR_CUDA(e->ffn_lo.need(cap * I * 2));
    // the GEMM reads whole 128-token tiles: rows past the live tokens must hold finite numbers
    for (DevBuf *b : {&e->hid_hi, &e->hid_lo, &e->ctx_hi, &e->ctx_lo, &e->ffn_hi, &e->ffn_lo})
        FR_CUDA(cudaMemsetAsync(b->p, 0, b->bytes, e->stream));
    const auto dt = CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
    using fr::mma::make_row_major_map;
    if (!make_row_major_map(&e->map_hid_hi, e->hid_hi.p, cap, dt, fr::enc::G_M, e->H) ||
        !make_row_major_map(&e->map_hid_lo, e->hid_lo.p, cap, dt, fr::enc::G_M, e->H) ||
        !make_row_major_map(&e->map_ctx_hi, e->ctx_hi.p, cap, dt, fr::enc::G_M, e->H) ||
        !make_row_major_map(&e->map_ctx_lo, e->ctx_lo.p, cap, dt, fr::enc::G_M, e->H) ||
        !make_row_major_map(&e->map_ffn_hi, e->ffn_hi.p, cap, dt, fr::enc::G_M, e->I) ||
        !make_row_major_map(&e->map_ffn_lo, e->ffn_lo.p, cap, dt, fr::enc::G_M, e->I))
        return fail(FR_ECUDA, "cuTensorMapEncodeTiled failed for the encoder activations");
    e->m_cap = cap;
    return FR_OK;
}

template <int EPI>
int gemm(fr_encoder *e, const CUtensorMap &xh, const CUtensorMap &xl, SplitW &w, int M, float *of32, bf16 *ohi, bf16 *olo,
         int pitch, cudaStream_t s) {
    auto kern = fr::enc::encoder_gemm_kernel<EPI>;
    FR_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(fr::enc::G_SMEM)));
    const dim3 grid((M + fr::enc::G_M - 1) / fr::enc::G_M, w.N / fr::enc::G_N);
    kern<<<grid, 192, fr::enc::G_SMEM, s>>>(xh, xl, w.map_hi, w.map_lo, M, w.K / fr::mma::K_CHUNK,
                                            static_cast<const float *>(w.bias.p), of32, ohi, olo, pitch);
    fr::count_launch();
    FR_CUDA(cudaGetLastError());
    (void)e;
    return FR_OK;
}

int forward_on_stream(fr_encoder *e, const int32_t *d_ids, const int32_t *d_lens, int B, int T, int pooling, int normalize,
                      float *d_out, float *d_hidden_out, cudaStream_t s) {
    const int M = B * T, H = e->H;
    int rc = ensure_activations(e, M);
    if (rc != FR_OK) return rc;
    float *hid = static_cast<float *>(e->hid.p), *tmp = static_cast<float *>(e->tmp.p), *qkv = static_cast<float *>(e->qkv.p);
    bf16 *hh = static_cast<bf16 *>(e->hid_hi.p), *hl = static_cast<bf16 *>(e->hid_lo.p);
    bf16 *ch = static_cast<bf16 *>(e->ctx_hi.p), *cl = static_cast<bf16 *>(e->ctx_lo.p);
    bf16 *fh = static_cast<bf16 *>(e->ffn_hi.p), *fl = static_cast<bf16 *>(e->ffn_lo.p);
    const int tok_grid = (M + 3) / 4;
    fr::enc::embed_ln_kernel<<<tok_grid, 128, 0, s>>>(d_ids, M, T, H, e->vocab, static_cast<const float *>(e->word.p),
                                                      static_cast<const float *>(e->pos.p), static_cast<const float *>(e->type.p),
                                                      static_cast<const float *>(e->emb_ln.gamma.p),
                                                      static_cast<const float *>(e->emb_ln.beta.p), e->eps, hid, hh, hl);
    fr::count_launch();
    FR_CUDA(cudaGetLastError());
    const int hd = H / e->heads;
    const size_t att_smem = (static_cast<size_t>(2) * T * (hd + 1) + static_cast<size_t>(4) * T) * sizeof(float);
    const float scale = 1.0f / sqrtf(static_cast<float>(hd));
    for (Layer &ly : e->layers) {
        rc = gemm<fr::enc::EPI_F32>(e, e->map_hid_hi, e->map_hid_lo, ly.qkv, M, qkv, nullptr, nullptr, 3 * H, s);
        if (rc != FR_OK) return rc;
        if (hd == 32)
            fr::enc::attention_kernel<32><<<dim3(B, e->heads), 128, att_smem, s>>>(qkv, d_lens, T, H, scale, ch, cl);
        else
            fr::enc::attention_kernel<64><<<dim3(B, e->heads), 128, att_smem, s>>>(qkv, d_lens, T, H, scale, ch, cl);
        fr::count_launch();
        FR_CUDA(cudaGetLastError());
        rc = gemm<fr::enc::EPI_F32>(e, e->map_ctx_hi, e->map_ctx_lo, ly.out, M, tmp, nullptr, nullptr, H, s);
        if (rc != FR_OK) return rc;
        fr::enc::add_ln_kernel<<<tok_grid, 128, 0, s>>>(tmp, M, H, static_cast<const float *>(ly.ln1.gamma.p),
                                                        static_cast<const float *>(ly.ln1.beta.p), e->eps, hid, hh, hl);
        fr::count_launch();
        FR_CUDA(cudaGetLastError());
        rc = gemm<fr::enc::EPI_GELU_SPLIT>(e, e->map_hid_hi, e->map_hid_lo, ly.ffn1, M, nullptr, fh, fl, e->I, s);
        if (rc != FR_OK) return rc;
        rc = gemm<fr::enc::EPI_F32>(e, e->map_ffn_hi, e->map_ffn_lo, ly.ffn2, M, tmp, nullptr, nullptr, H, s);
        if (rc != FR_OK) return rc;
        fr::enc::add_ln_kernel<<<tok_grid, 128, 0, s>>>(tmp, M, H, static_cast<const float *>(ly.ln2.gamma.p),
                                                        static_cast<const float *>(ly.ln2.beta.p), e->eps, hid, hh, hl);
        fr::count_launch();
        FR_CUDA(cudaGetLastError());
    }
    if (d_hidden_out)
        FR_CUDA(cudaMemcpyAsync(d_hidden_out, hid, static_cast<size_t>(M) * H * sizeof(float), cudaMemcpyDeviceToDevice, s));
    fr::enc::pool_kernel<<<B, 128, 0, s>>>(hid, d_lens, T, H, pooling, normalize, d_out);
    fr::count_launch();
    FR_CUDA(cudaGetLastError());
    return FR_OK;
}

int check_forward_args(fr_encoder *e, const void *ids, const void *lens, int B, int T, int pooling, const void *out) {
    if (!e) return fail(FR_EINVAL, "encoder is NULL");
    if (!e->ready) return fail(FR_EINVAL, "encoder weights are incomplete: call fr_encoder_finalize (it names what is missing)");
    if (B < 0 || T < 1) return fail(FR_EINVAL, "bad batch / sequence length (%d / %d)", B, T);
    if (T > e->max_pos) return fail(FR_EINVAL, "sequence length %d exceeds max_position_embeddings %d", T, e->max_pos);
    if (pooling != 0 && pooling != 1) return fail(FR_EINVAL, "pooling must be 0 (CLS) or 1 (mean)");
    if (B > 0 && (!ids || !lens || !out)) return fail(FR_EINVAL, "NULL buffer");
    return FR_OK;
}

}  // namespace

extern "C" {

int fr_encoder_create(int device, int vocab_size, int hidden_size, int num_layers, int num_heads, int intermediate_size,
                      int max_position_embeddings, int type_vocab_size, float layer_norm_eps, fr_encoder **out) {
    if (!out) return fail(FR_EINVAL, "out is NULL");
    *out = nullptr;
    if (vocab_size < 1 || num_layers < 1 || num_heads < 1 || max_position_embeddings < 1 || type_vocab_size < 1)
        return fail(FR_EINVAL, "bad model dimensions");
    if (hidden_size % 128 != 0 || hidden_size > 1024 || intermediate_size % 128 != 0)
        return fail(FR_EUNSUP, "hidden_size must be a multiple of 128 up to 1024 and intermediate_size a multiple of 128 "
                               "(got %d / %d)", hidden_size, intermediate_size);
    if (hidden_size % num_heads != 0 || (hidden_size / num_heads != 32 && hidden_size / num_heads != 64))
        return fail(FR_EUNSUP, "head size must be 32 or 64 (got %d / %d)", hidden_size, num_heads);
    if (max_position_embeddings > 512) return fail(FR_EUNSUP, "sequences up to 512 tokens (got %d)", max_position_embeddings);
    int rc = fr::check_device(device, nullptr);
    if (rc != FR_OK) return rc;
    DeviceGuard g(device);
    fr_encoder *e = new (std::nothrow) fr_encoder();
    if (!e) return fail(FR_ENOMEM, "host allocation failed");
    e->device = device;
    e->vocab = vocab_size;
    e->H = hidden_size;
    e->L = num_layers;
    e->heads = num_heads;
    e->I = intermediate_size;
    e->max_pos = max_position_embeddings;
    e->type_vocab = type_vocab_size;
    e->eps = layer_norm_eps;
    e->layers.resize(static_cast<size_t>(num_layers));
    if (cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking) != cudaSuccess) {
        delete e;
        return fail(FR_ECUDA, "stream creation failed");
    }
    const int hd = hidden_size / num_heads;
    const size_t att_smem = (static_cast<size_t>(2) * max_position_embeddings * (hd + 1) + static_cast<size_t>(4) * max_position_embeddings) * 4;
    cudaError_t ce = hd == 32 ? cudaFuncSetAttribute(fr::enc::attention_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(att_smem))
                              : cudaFuncSetAttribute(fr::enc::attention_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(att_smem));
    if (ce != cudaSuccess) {
        fr_encoder_destroy(e);
        return fail(FR_ECUDA, "attention kernel needs %zu bytes of shared memory: %s", att_smem, cudaGetErrorString(ce));
    }
    *out = e;
    return FR_OK;
}

int fr_encoder_destroy(fr_encoder *e) {
    if (!e) return FR_OK;
    {
        DeviceGuard g(e->device);
        cudaDeviceSynchronize();
        for (DevBuf *b : {&e->word, &e->pos, &e->type, &e->emb_ln.gamma, &e->emb_ln.beta, &e->hid, &e->hid_hi, &e->hid_lo,
                          &e->qkv, &e->ctx_hi, &e->ctx_lo, &e->tmp, &e->ffn_hi, &e->ffn_lo, &e->ids, &e->lens, &e->out, &e->stage})
            b->release();
        for (Layer &ly : e->layers) {
            for (SplitW *w : {&ly.qkv, &ly.out, &ly.ffn1, &ly.ffn2}) {
                w->hi.release();
                w->lo.release();
                w->bias.release();
            }
            for (LnP *p : {&ly.ln1, &ly.ln2}) {
                p->gamma.release();
                p->beta.release();
            }
        }
        e->pin.release();
        for (auto &kv : e->graphs)
            if (kv.second.exec) cudaGraphExecDestroy(kv.second.exec);
        if (e->stream) cudaStreamDestroy(e->stream);
    }
    delete e;
    return FR_OK;
}

// Names are the keys of a Hugging Face BertModel state dict (with or without a leading "bert."):
//   embeddings.{word,position,token_type}_embeddings.weight, embeddings.LayerNorm.{weight,bias},
//   encoder.layer.<i>.attention.self.{query,key,value}.{weight,bias}, encoder.layer.<i>.attention.output.dense.{weight,bias},
//   encoder.layer.<i>.attention.output.LayerNorm.{weight,bias}, encoder.layer.<i>.intermediate.dense.{weight,bias},
//   encoder.layer.<i>.output.dense.{weight,bias}, encoder.layer.<i>.output.LayerNorm.{weight,bias}.
// Unknown names (pooler.*, position_ids) are ignored and reported as such through *out_used = 0.
int fr_encoder_set_tensor(fr_encoder *e, const char *name, const float *data, int64_t n, int *out_used) {
    if (!e || !name || !data) return fail(FR_EINVAL, "NULL argument");
    if (out_used) *out_used = 0;
    std::lock_guard<std::mutex> lk(e->mu);
    DeviceGuard g(e->device);
    std::string s(name);
    if (s.rfind("bert.", 0) == 0) s = s.substr(5);
    const int H = e->H, I = e->I;
    auto want = [&](int64_t expect) -> int {
        if (n != expect) return fail(FR_EINVAL, "tensor '%s' has %lld elements, the model needs %lld", name, (long long)n, (long long)expect);
        return FR_OK;
    };
    int rc = FR_OK;
    e->ready = false;
    if (s == "embeddings.word_embeddings.weight") {
        if ((rc = want(static_cast<int64_t>(e->vocab) * H)) != FR_OK) return rc;
        rc = upload(e->word, data, n, e->stream);
        e->have_word = rc == FR_OK;
    } else if (s == "embeddings.position_embeddings.weight") {
        if ((rc = want(static_cast<int64_t>(e->max_pos) * H)) != FR_OK) return rc;
        rc = upload(e->pos, data, n, e->stream);
        e->have_pos = rc == FR_OK;
    } else if (s == "embeddings.token_type_embeddings.weight") {
        if ((rc = want(static_cast<int64_t>(e->type_vocab) * H)) != FR_OK) return rc;
        rc = upload(e->type, data, H, e->stream);  // queries are single-segment: only type 0 is ever read
        e->have_type = rc == FR_OK;
    } else if (s == "embeddings.LayerNorm.weight" || s == "embeddings.LayerNorm.bias") {
        if ((rc = want(H)) != FR_OK) return rc;
        const bool w = s.back() == 't';
        rc = upload(w ? e->emb_ln.gamma : e->emb_ln.beta, data, n, e->stream);
        (w ? e->emb_ln.have_g : e->emb_ln.have_b) = rc == FR_OK;
    } else if (s.rfind("encoder.layer.", 0) == 0) {
        const size_t p0 = 14, p1 = s.find('.', p0);
        if (p1 == std::string::npos) return FR_OK;
        const int li = std::atoi(s.substr(p0, p1 - p0).c_str());
        if (li < 0 || li >= e->L) return fail(FR_EINVAL, "tensor '%s': the model has %d layers", name, e->L);
        Layer &ly = e->layers[static_cast<size_t>(li)];
        const std::string t = s.substr(p1 + 1);
        const bool is_w = t.size() > 7 && t.compare(t.size() - 7, 7, ".weight") == 0;
        auto lin = [&](SplitW &w, int N, int K, int row0, int rows) -> int {
            if (is_w) {
                if ((rc = want(static_cast<int64_t>(rows) * K)) != FR_OK) return rc;
                return upload_split_rows(e, w, N, K, row0, rows, data);
            }
            if ((rc = want(rows)) != FR_OK) return rc;
            return upload_bias_rows(e, w, N, row0, rows, data);
        };
        auto lnp = [&](LnP &p) -> int {
            if ((rc = want(H)) != FR_OK) return rc;
            rc = upload(is_w ? p.gamma : p.beta, data, n, e->stream);
            (is_w ? p.have_g : p.have_b) = rc == FR_OK;
            return rc;
        };
        const std::string stem = t.substr(0, t.rfind('.'));
        if (stem == "attention.self.query" || stem == "attention.self.key" || stem == "attention.self.value") {
            const int part = stem.back() == 'y' ? (stem[15] == 'q' ? 0 : 1) : 2;  // quer-y, ke-y, valu-e
            rc = lin(ly.qkv, 3 * H, H, part * H, H);
            if (rc == FR_OK) (is_w ? ly.qkv_parts : ly.qkv_bias_parts) |= 1 << part;
            ly.qkv.have_w = ly.qkv_parts == 7;
            ly.qkv.have_b = ly.qkv_bias_parts == 7;
        } else if (stem == "attention.output.dense") {
            rc = lin(ly.out, H, H, 0, H);
            (is_w ? ly.out.have_w : ly.out.have_b) = rc == FR_OK;
        } else if (stem == "attention.output.LayerNorm") {
            rc = lnp(ly.ln1);
        } else if (stem == "intermediate.dense") {
            rc = lin(ly.ffn1, I, H, 0, I);
            (is_w ? ly.ffn1.have_w : ly.ffn1.have_b) = rc == FR_OK;
        } else if (stem == "output.dense") {
            rc = lin(ly.ffn2, H, I, 0, H);
            (is_w ? ly.ffn2.have_w : ly.ffn2.have_b) = rc == FR_OK;
        } else if (stem == "output.LayerNorm") {
            rc = lnp(ly.ln2);
        } else {
            return FR_OK;
        }
    } else {
        return FR_OK;  // pooler.*, embeddings.position_ids, ...
    }
    if (rc == FR_OK && out_used) *out_used = 1;
    return rc;
}

int fr_encoder_finalize(fr_encoder *e) {
    if (!e) return fail(FR_EINVAL, "encoder is NULL");
    std::lock_guard<std::mutex> lk(e->mu);
    DeviceGuard g(e->device);
    std::string missing;
    auto need = [&](bool have, const std::string &what) {
        if (!have && missing.size() < 300) missing += (missing.empty() ? "" : ", ") + what;
    };
    need(e->have_word, "embeddings.word_embeddings.weight");
    need(e->have_pos, "embeddings.position_embeddings.weight");
    need(e->have_type, "embeddings.token_type_embeddings.weight");
    need(e->emb_ln.have_g && e->emb_ln.have_b, "embeddings.LayerNorm");
    for (int i = 0; i < e->L; ++i) {
        Layer &ly = e->layers[static_cast<size_t>(i)];
        const std::string p = "encoder.layer." + std::to_string(i) + ".";
        need(ly.qkv.have_w && ly.qkv.have_b, p + "attention.self.{query,key,value}");
        need(ly.out.have_w && ly.out.have_b, p + "attention.output.dense");
        need(ly.ln1.have_g && ly.ln1.have_b, p + "attention.output.LayerNorm");
        need(ly.ffn1.have_w && ly.ffn1.have_b, p + "intermediate.dense");
        need(ly.ffn2.have_w && ly.ffn2.have_b, p + "output.dense");
        need(ly.ln2.have_g && ly.ln2.have_b, p + "output.LayerNorm");
    }
    if (!missing.empty()) return fail(FR_EINVAL, "encoder weights missing: %s", missing.c_str());
    for (Layer &ly : e->layers)
        for (SplitW *w : {&ly.qkv, &ly.out, &ly.ffn1, &ly.ffn2})
            if (!make_maps(*w)) return fail(FR_ECUDA, "cuTensorMapEncodeTiled failed for an encoder weight");
    for (auto &kv : e->graphs)  // graphs captured over an earlier set of weights bake their addresses in
        if (kv.second.exec) cudaGraphExecDestroy(kv.second.exec);
    e->graphs.clear();
    e->ready = true;
    return FR_OK;
}

int fr_encoder_forward_device(fr_encoder *e, const int32_t *d_ids, const int32_t *d_lens, int B, int T, int pooling,
                              int normalize, float *d_out, float *d_hidden_or_null, void *stream) {
    int rc = check_forward_args(e, d_ids, d_lens, B, T, pooling, d_out);
    if (rc != FR_OK || B == 0) return rc;
    std::lock_guard<std::mutex> lk(e->mu);
    DeviceGuard g(e->device);
    return forward_on_stream(e, d_ids, d_lens, B, T, pooling, normalize, d_out, d_hidden_or_null, static_cast<cudaStream_t>(stream));
}

int fr_encoder_forward(fr_encoder *e, const int32_t *ids, const int32_t *lens, int B, int T, int pooling, int normalize,
                       float *out, float *hidden_or_null) {
    int rc = check_forward_args(e, ids, lens, B, T, pooling, out);
    if (rc != FR_OK || B == 0) return rc;
    std::lock_guard<std::mutex> lk(e->mu);
    DeviceGuard g(e->device);
    const size_t M = static_cast<size_t>(B) * T, H = e->H;
    cudaStream_t s = e->stream;
    const size_t ib = (M * 4 + 15) & ~size_t(15), lb = (static_cast<size_t>(B) * 4 + 15) & ~size_t(15), ob = static_cast<size_t>(B) * H * 4;
    FR_CUDA(e->ids.need(M * 4));
    FR_CUDA(e->lens.need(static_cast<size_t>(B) * 4));
    FR_CUDA(e->out.need(ob));
    FR_CUDA(e->pin.need(ib + lb + ob));
    uint8_t *pin = static_cast<uint8_t *>(e->pin.p);
    std::memcpy(pin, ids, M * 4);
    std::memcpy(pin + ib, lens, static_cast<size_t>(B) * 4);
    auto enqueue = [&]() -> int {
        FR_CUDA(cudaMemcpyAsync(e->ids.p, pin, M * 4, cudaMemcpyHostToDevice, s));
        FR_CUDA(cudaMemcpyAsync(e->lens.p, pin + ib, static_cast<size_t>(B) * 4, cudaMemcpyHostToDevice, s));
        int r = forward_on_stream(e, static_cast<const int32_t *>(e->ids.p), static_cast<const int32_t *>(e->lens.p), B, T,
                                  pooling, normalize, static_cast<float *>(e->out.p), nullptr, s);
        if (r != FR_OK) return r;
        FR_CUDA(cudaMemcpyAsync(pin + ib + lb, e->out.p, ob, cudaMemcpyDeviceToHost, s));
        return FR_OK;
    };
    auto state = [&]() {  // every address a captured graph bakes in
        uint64_t h = 1469598103934665603ull;
        for (const void *p : {e->pin.p, e->ids.p, e->lens.p, e->out.p, e->hid.p, e->hid_hi.p, e->hid_lo.p, e->qkv.p, e->ctx_hi.p,
                              e->ctx_lo.p, e->tmp.p, e->ffn_hi.p, e->ffn_lo.p}) {
            h ^= reinterpret_cast<uintptr_t>(p);
            h *= 1099511628211ull;
        }
        return h;
    };
    fr_encoder::Graph *gr = nullptr;
    // graph work only while no caller of the library waits for results (fr_host.h: graph_wait_mutex); eager otherwise
    std::unique_lock<std::shared_mutex> graph_lk(fr::graph_wait_mutex(), std::try_to_lock);
    if (graph_lk.owns_lock() && !hidden_or_null && static_cast<int64_t>(M) <= e->m_cap) {  // (the first call of a size runs eagerly: it grows the buffers)
        if (e->graphs.size() > 64) {
            for (auto &kv : e->graphs)
                if (kv.second.exec) cudaGraphExecDestroy(kv.second.exec);
            e->graphs.clear();
        }
        gr = &e->graphs[(static_cast<uint64_t>(B) << 32) | (static_cast<uint64_t>(T) << 8) | (static_cast<uint64_t>(pooling) << 1) |
                        static_cast<uint64_t>(normalize ? 1 : 0)];
        const uint64_t h = state();
        if (gr->exec && gr->state != h) {
            cudaGraphExecDestroy(gr->exec);
            gr->exec = nullptr;
        }
        if (!gr->exec && !gr->failed && gr->seen == h) {
            const int64_t l0 = fr_launch_count();
            cudaGraph_t graph = nullptr;
            bool ok = cudaStreamBeginCapture(s, cudaStreamCaptureModeRelaxed) == cudaSuccess;
            if (ok) {
                const int r = enqueue();
                ok = cudaStreamEndCapture(s, &graph) == cudaSuccess && r == FR_OK && graph != nullptr;
            }
            gr->launches = fr_launch_count() - l0;
            fr::count_launches(-gr->launches);  // the capture enqueued nothing
            if (ok) ok = state() == h && cudaGraphInstantiate(&gr->exec, graph, 0) == cudaSuccess;
            if (graph) cudaGraphDestroy(graph);
            if (!ok) {
                cudaGetLastError();
                gr->exec = nullptr;
                gr->failed = true;
            } else {
                gr->state = h;
            }
        }
    }
    if (gr && gr->exec) {
        FR_CUDA(cudaGraphLaunch(gr->exec, s));
        fr::count_launches(gr->launches);
        e->n_graph_replays += 1;
    } else {
        rc = enqueue();
        if (rc != FR_OK) {
            cudaStreamSynchronize(s);
            return rc;
        }
        if (gr) gr->seen = state();
        if (hidden_or_null) FR_CUDA(cudaMemcpyAsync(hidden_or_null, e->hid.p, M * H * 4, cudaMemcpyDeviceToHost, s));
    }
    if (graph_lk.owns_lock()) graph_lk.unlock();
    {
        std::shared_lock<std::shared_mutex> wait_lk(fr::graph_wait_mutex());
        FR_CUDA(cudaStreamSynchronize(s));
    }
    std::memcpy(out, pin + ib + lb, ob);
    return FR_OK;
}

}  // extern "C"
