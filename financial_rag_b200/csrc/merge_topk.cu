// K3 merge_topk -- merge sorted partial top-k lists into the final sorted top-k (sm_100a).
//
// Two uses on the path that replaces chromadb Collection.query
// (parent_child/chroma_child_store.py:63):
//   LOCAL : the per-CTA lists K1/K2 wrote for one shard  -> that shard's result
//   SHARDS: the G per-GPU lists brought together by the NCCL all-gather (K4) -> the final result,
//           ties resolved by (shard, position) == global insertion order, so any G gives the
//           same answer as G = 1.
// One CTA per query; each warp folds a slice of the lists into a register WarpTopK using the
// threshold gate, warps are merged through shared memory, warp 0 writes the result and turns
// packed keys into (distance, int64 key).  Latency-bound; traffic is P*k*8 bytes per query.
#include "fr_common.cuh"
#include "fr_kernels.h"

namespace fr {

// SHARDS mode, two tie orders (a tie = equal score bits):
//   block layout  (cyclic_w == 0): shard g holds a contiguous block of global rows, so the global insertion order of
//                 tied candidates is (shard, position in the shard's sorted list): low word = ~(g * k + position);
//   cyclic layout (cyclic_w == W): global row s lives on shard s % W at local row s / W (fr_group: a collection that
//                 grows by upserts cannot know its final block boundaries), so the global insertion order is
//                 local_row * W + shard -- that IS the global row, and it goes into the low word (fits 32 bits:
//                 fr_group caps a shard at (2^32 - 16) / W rows).  The position needed to fetch the int64 key is
//                 recovered by a binary search of the shard's (strictly descending) list.
__device__ __forceinline__ uint64_t shard_order_key(uint64_t key, int p, int pos, int k, int cyclic_w) {
    const uint32_t ord = cyclic_w > 0 ? key_row(key) * static_cast<uint32_t>(cyclic_w) + static_cast<uint32_t>(p)
                                      : static_cast<uint32_t>(p * k + pos);
    return (key & 0xffffffff00000000ull) | static_cast<uint64_t>(0xffffffffu - ord);
}
__device__ __forceinline__ int64_t shard_key_of(uint64_t key, const uint64_t *__restrict__ packed,
                                                const int64_t *__restrict__ shard_keys, int64_t shard_stride, int b,
                                                int k, int cyclic_w) {
    const uint32_t idx = key_row(key);
    if (cyclic_w == 0) {
        const int g = idx / k, jj = idx - g * k;
        return shard_keys[static_cast<int64_t>(g) * shard_stride + static_cast<int64_t>(b) * k + jj];
    }
    const int g = idx % static_cast<uint32_t>(cyclic_w);
    const uint32_t local_row = idx / static_cast<uint32_t>(cyclic_w);
    const uint64_t orig = (key & 0xffffffff00000000ull) | static_cast<uint64_t>(0xffffffffu - local_row);
    const int64_t base = static_cast<int64_t>(g) * shard_stride + static_cast<int64_t>(b) * k;
    int lo = 0, hi = k - 1;  // descending list, `orig` is in it
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (packed[base + mid] > orig) lo = mid + 1; else hi = mid;
    }
    return shard_keys[base + lo];
}

template <int KPL, bool SHARDS>
__global__ void __launch_bounds__(128)
merge_topk_kernel(const uint64_t *__restrict__ packed, int P, int64_t shard_stride, int B, int k,
                  const int64_t *__restrict__ row_keys, const int64_t *__restrict__ shard_keys, bool l2,
                  float *__restrict__ out_dist, uint64_t *__restrict__ out_packed,
                  int64_t *__restrict__ out_keys, const uint8_t *__restrict__ only_flagged,
                  const int *__restrict__ limit, int cyclic_w) {
    __shared__ uint64_t lists[4 * 32 * KPL];
    const int b = blockIdx.x;
    if (only_flagged != nullptr && only_flagged[b] == 0) return;  // uniform per CTA
    if (limit != nullptr && b >= *limit) return;
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int nwarps = blockDim.x >> 5;

    WarpTopK<KPL> tk[1];
    tk[0].clear();
    uint64_t thr = 0ull;
    for (int p = warp; p < P; p += nwarps) {
        const uint64_t *src = packed + static_cast<int64_t>(p) * shard_stride + static_cast<int64_t>(b) * k;
        for (int i0 = 0; i0 < k; i0 += 32) {
            const int i = i0 + lane;
            uint64_t key = (i < k) ? src[i] : 0ull;
            if (SHARDS && key != 0ull) key = shard_order_key(key, p, i, k, cyclic_w);
            unsigned m = __ballot_sync(FULL_MASK, key != 0ull && key > thr);
            if (m == 0u) break;  // list is sorted: nothing further in it can enter
            while (m) {
                const int srcl = __ffs(m) - 1;
                m &= m - 1;
                const uint64_t kk = __shfl_sync(FULL_MASK, key, srcl);
                tk[0].insert(kk, k, lane);
            }
            thr = tk[0].kth(k);
        }
    }
    cta_merge_lists<KPL, 1>(tk, lists, nwarps, warp, lane, k);
    if (warp != 0) return;
#pragma unroll
    for (int j = 0; j < KPL; ++j) {
        const int i = j * 32 + lane;
        if (i >= k) continue;
        const uint64_t key = tk[0].e[j];
        const int64_t o = static_cast<int64_t>(b) * k + i;
        if (key == 0ull) {
            if (out_dist) out_dist[o] = INFINITY;
            if (out_packed) out_packed[o] = 0ull;
            out_keys[o] = -1;
            continue;
        }
        const uint32_t idx = key_row(key);
        const float s = key_score(key);
        if (out_dist) out_dist[o] = l2 ? -s : 1.0f - s;
        if (out_packed) out_packed[o] = key;
        if (SHARDS) {
            out_keys[o] = shard_key_of(key, packed, shard_keys, shard_stride, b, k, cyclic_w);
        } else {
            out_keys[o] = row_keys[idx];
        }
    }
}

// k <= 32: every partial list is one sorted 32-entry vector (zero padded), so folding a list into the running
// top-32 is an elementwise max against the reversed list + one 5-stage bitonic merge -- no per-entry inserts.
// 8 warps fold P/8 lists each (next list prefetched while the current one is merged), then a 3-level tree.
template <bool SHARDS>
__global__ void __launch_bounds__(256)
merge_topk32_kernel(const uint64_t *__restrict__ packed, int P, int64_t shard_stride, int B, int k,
                    const int64_t *__restrict__ row_keys, const int64_t *__restrict__ shard_keys, bool l2,
                    float *__restrict__ out_dist, uint64_t *__restrict__ out_packed,
                    int64_t *__restrict__ out_keys, const uint8_t *__restrict__ only_flagged,
                    const int *__restrict__ limit, int cyclic_w) {
    __shared__ uint64_t lists[8][32];
    const int b = blockIdx.x;
    if (only_flagged != nullptr && only_flagged[b] == 0) return;  // uniform per CTA
    if (limit != nullptr && b >= *limit) return;
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    auto load = [&](int p) -> uint64_t {
        if (p >= P || lane >= k) return 0ull;
        uint64_t key = packed[static_cast<int64_t>(p) * shard_stride + static_cast<int64_t>(b) * k + lane];
        if (SHARDS && key != 0ull) key = shard_order_key(key, p, lane, k, cyclic_w);
        return key;
    };
    uint64_t run = 0ull;
    uint64_t next = load(warp);
    for (int p = warp; p < P; p += 8) {
        const uint64_t cur = next;
        next = load(p + 8);
        run = bitonic_merge32_desc(umax64(run, reverse32(cur, lane)), lane);
    }
    lists[warp][lane] = run;
    for (int stride = 4; stride > 0; stride >>= 1) {
        __syncthreads();
        if (warp < stride) {
            run = bitonic_merge32_desc(umax64(run, reverse32(lists[warp + stride][lane], lane)), lane);
            lists[warp][lane] = run;
        }
    }
    if (warp != 0 || lane >= k) return;
    const int64_t o = static_cast<int64_t>(b) * k + lane;
    if (run == 0ull) {
        if (out_dist) out_dist[o] = INFINITY;
        if (out_packed) out_packed[o] = 0ull;
        out_keys[o] = -1;
        return;
    }
    const uint32_t idx = key_row(run);
    const float sc = key_score(run);
    if (out_dist) out_dist[o] = l2 ? -sc : 1.0f - sc;
    if (out_packed) out_packed[o] = run;
    if (SHARDS) {
        out_keys[o] = shard_key_of(run, packed, shard_keys, shard_stride, b, k, cyclic_w);
    } else {
        out_keys[o] = row_keys[idx];
    }
}

cudaError_t launch_merge_topk(const MergeArgs &a) {
    if (a.B <= 0) return cudaSuccess;
    const dim3 grid(a.B), block(128);
#define FR_MERGE(KPL, SH)                                                                          \
    merge_topk_kernel<KPL, SH><<<grid, block, 0, a.stream>>>(a.packed, a.P, a.shard_stride, a.B, a.k, \
                                                             a.row_keys, a.shard_keys, a.l2, a.out_dist, \
                                                             a.out_packed, a.out_keys, a.only_flagged, a.limit, a.cyclic_world)
    if (a.k <= 32) {
        if (a.shards)
            merge_topk32_kernel<true><<<grid, 256, 0, a.stream>>>(a.packed, a.P, a.shard_stride, a.B, a.k, a.row_keys,
                                                                 a.shard_keys, a.l2, a.out_dist, a.out_packed, a.out_keys,
                                                                 a.only_flagged, a.limit, a.cyclic_world);
        else
            merge_topk32_kernel<false><<<grid, 256, 0, a.stream>>>(a.packed, a.P, a.shard_stride, a.B, a.k, a.row_keys,
                                                                  a.shard_keys, a.l2, a.out_dist, a.out_packed, a.out_keys,
                                                                  a.only_flagged, a.limit, a.cyclic_world);
    } else if (a.k <= 128) {
        if (a.shards) FR_MERGE(4, true); else FR_MERGE(4, false);
    } else if (a.k <= 256 && !a.shards) {
        FR_MERGE(8, false);  // K2 second-chance lists
    } else {
        return cudaErrorInvalidValue;
    }
#undef FR_MERGE
    count_launch();
    return cudaGetLastError();
}

}  // namespace fr
