// K0 normalize_cast_append -- the device side of Collection.upsert
// (parent_child/chroma_child_store.py:54-59).
//
// fp32 rows arrive from the embedder (parent_child/pipeline.py:140-143).  Cosine collections store
// x * 1/(||x|| + 1e-30) (hnswlib's normalisation, fp32 arithmetic); the row is then rounded to
// bf16 (RNE) or kept as fp32 and written to its slot of the shard; the int64 key goes to keys[].
// One warp per vector; HBM-bound (reads 4*dim, writes 2*dim or 4*dim bytes per row).
#include "fr_common.cuh"
#include "fr_kernels.h"

namespace fr {

__global__ void __launch_bounds__(256)
ingest_kernel(const float *__restrict__ src, const int64_t *__restrict__ src_keys, int64_t first_key,
              const int64_t *__restrict__ target_rows, int64_t base_row, int64_t n, int dim, bool normalize,
              bool bf16, uint8_t *__restrict__ corpus, int64_t *__restrict__ keys) {
    const int lane = threadIdx.x & 31;
    const int64_t gwarp = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    const int64_t gstride = (static_cast<int64_t>(gridDim.x) * blockDim.x) >> 5;
    const int c4 = dim >> 2;
    for (int64_t i = gwarp; i < n; i += gstride) {
        const int64_t t = target_rows ? target_rows[i] : base_row + i;
        if (t < 0) continue;  // superseded by a later vector with the same key in this call
        const float4 *s4 = reinterpret_cast<const float4 *>(src + i * dim);
        float inv = 1.0f;
        if (normalize) {
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
            inv = 1.0f / (sqrtf(ss) + 1e-30f);
        }
        if (bf16) {
            uint2 *d = reinterpret_cast<uint2 *>(corpus + static_cast<size_t>(t) * dim * 2);
            for (int c = lane; c < c4; c += 32) {
                const float4 v = s4[c];
                const __nv_bfloat162 lo = __floats2bfloat162_rn(v.x * inv, v.y * inv);
                const __nv_bfloat162 hi = __floats2bfloat162_rn(v.z * inv, v.w * inv);
                uint2 o;
                o.x = *reinterpret_cast<const uint32_t *>(&lo);
                o.y = *reinterpret_cast<const uint32_t *>(&hi);
                d[c] = o;
            }
        } else {
            float4 *d = reinterpret_cast<float4 *>(corpus + static_cast<size_t>(t) * dim * 4);
            for (int c = lane; c < c4; c += 32) {
                float4 v = s4[c];
                v.x *= inv; v.y *= inv; v.z *= inv; v.w *= inv;
                d[c] = v;
            }
        }
        if (lane == 0) keys[t] = src_keys ? src_keys[i] : first_key + i;
    }
}

__global__ void fill_keys_kernel(int64_t *keys, const int64_t *rows, int64_t n, int64_t value) {
    const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i < n) keys[rows[i]] = value;
}

cudaError_t launch_ingest(const IngestArgs &a) {
    if (a.n <= 0) return cudaSuccess;
    int64_t blocks = (a.n + 7) / 8;  // 8 warps per CTA, one vector per warp per iteration
    if (blocks > 148 * 16) blocks = 148 * 16;
    ingest_kernel<<<static_cast<int>(blocks), 256, 0, a.stream>>>(a.src, a.src_keys, a.first_key, a.target_rows,
                                                                  a.base_row, a.n, a.dim, a.normalize, a.bf16,
                                                                  a.corpus, a.keys);
    count_launch();
    return cudaGetLastError();
}

// fp32 rows -> bf16 (RNE), 8 elements per thread: the selection copy ("shadow") of an fp32 collection
__global__ void shadow_convert_kernel(const float4 *__restrict__ src, uint4 *__restrict__ dst, int64_t n8) {
    const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= n8) return;
    const float4 a = src[2 * i], b = src[2 * i + 1];
    const __nv_bfloat162 p0 = __floats2bfloat162_rn(a.x, a.y), p1 = __floats2bfloat162_rn(a.z, a.w);
    const __nv_bfloat162 p2 = __floats2bfloat162_rn(b.x, b.y), p3 = __floats2bfloat162_rn(b.z, b.w);
    uint4 o;
    o.x = *reinterpret_cast<const uint32_t *>(&p0);
    o.y = *reinterpret_cast<const uint32_t *>(&p1);
    o.z = *reinterpret_cast<const uint32_t *>(&p2);
    o.w = *reinterpret_cast<const uint32_t *>(&p3);
    dst[i] = o;
}

cudaError_t launch_shadow_convert(const float *src, void *dst_bf16, int64_t n_elems, cudaStream_t s) {
    if (n_elems <= 0) return cudaSuccess;
    if (n_elems % 8 != 0) return cudaErrorInvalidValue;
    const int64_t n8 = n_elems / 8;
    shadow_convert_kernel<<<static_cast<unsigned>((n8 + 255) / 256), 256, 0, s>>>(reinterpret_cast<const float4 *>(src),
                                                                               static_cast<uint4 *>(dst_bf16), n8);
    count_launch();
    return cudaGetLastError();
}

cudaError_t launch_fill_keys(int64_t *keys, const int64_t *rows, int64_t n, int64_t value, cudaStream_t s) {
    if (n <= 0) return cudaSuccess;
    fill_keys_kernel<<<static_cast<int>((n + 255) / 256), 256, 0, s>>>(keys, rows, n, value);
    count_launch();
    return cudaGetLastError();
}

}  // namespace fr
