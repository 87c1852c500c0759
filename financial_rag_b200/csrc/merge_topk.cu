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

// k in (32, 256]: a partial list is 32-entry blocks sorted descending, block after block.  Folding a block into the
// running top-(32 KPL) list (registers: entry i in lane i & 31, slot i >> 5) is one bitonic network (fold_sorted32), and a
// list is abandoned at the first block whose best entry does not beat the running list's last one.  8 warps fold P / 8
// lists each, then a 3-level tree through shared memory.  (The first version inserted candidates one at a time:
// 178 us for ONE query's 148 lists of 128 -- cfg3's top-50 -- against ~15 us now; profiles/r02_launches_cfg3_b64.txt.)
template <int KPL>
__device__ __forceinline__ uint64_t wide_last(const uint64_t (&R)[KPL]) {
    const uint32_t lo = __shfl_sync(FULL_MASK, static_cast<uint32_t>(R[KPL - 1]), 31);
    const uint32_t hi = __shfl_sync(FULL_MASK, static_cast<uint32_t>(R[KPL - 1] >> 32), 31);
    return (static_cast<uint64_t>(hi) << 32) | lo;
}
// folds one sorted list (`n` entries, `load(i)` yields entry i or 0; `first` = its block 0, loaded by the caller ahead
// of time so that the next list's load overlaps this list's networks) into R
template <int KPL, typename Load>
__device__ __forceinline__ void wide_fold_list(uint64_t (&R)[KPL], int n, int lane, uint64_t first, Load load) {
    for (int i0 = 0; i0 < n; i0 += 32) {
        const uint64_t key = i0 == 0 ? first : load(i0 + lane);
        const uint32_t hlo = __shfl_sync(FULL_MASK, static_cast<uint32_t>(key), 0);
        const uint32_t hhi = __shfl_sync(FULL_MASK, static_cast<uint32_t>(key >> 32), 0);
        const uint64_t head = (static_cast<uint64_t>(hhi) << 32) | hlo;
        if (head == 0ull || head <= wide_last<KPL>(R)) break;  // sorted: nothing further in this list can enter
        fold_sorted32<KPL>(R, key, lane);
    }
}

template <int KPL, bool SHARDS>
__global__ void __launch_bounds__(256)
merge_topk_wide_kernel(const uint64_t *__restrict__ packed, int P, int64_t shard_stride, int B, int k,
                       const int64_t *__restrict__ row_keys, const int64_t *__restrict__ shard_keys, bool l2,
                       float *__restrict__ out_dist, uint64_t *__restrict__ out_packed,
                       int64_t *__restrict__ out_keys, const uint8_t *__restrict__ only_flagged,
                       const int *__restrict__ limit, int cyclic_w) {
    constexpr int CAP = 32 * KPL;
    __shared__ uint64_t lists[8][CAP];
    const int b = blockIdx.x;
    if (only_flagged != nullptr && only_flagged[b] == 0) return;  // uniform per CTA
    if (limit != nullptr && b >= *limit) return;
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    uint64_t R[KPL];
#pragma unroll
    for (int j = 0; j < KPL; ++j) R[j] = 0ull;
    auto entry = [&](int p, int i) -> uint64_t {
        if (p >= P || i >= k) return 0ull;
        uint64_t key = packed[static_cast<int64_t>(p) * shard_stride + static_cast<int64_t>(b) * k + i];
        if (SHARDS && key != 0ull) key = shard_order_key(key, p, i, k, cyclic_w);
        return key;
    };
    uint64_t next = entry(warp, lane);
    for (int p = warp; p < P; p += 8) {
        const uint64_t cur = next;
        next = entry(p + 8, lane);
        wide_fold_list<KPL>(R, k, lane, cur, [&](int i) -> uint64_t { return entry(p, i); });
    }
#pragma unroll
    for (int j = 0; j < KPL; ++j) lists[warp][j * 32 + lane] = R[j];
    for (int stride = 4; stride > 0; stride >>= 1) {
        __syncthreads();
        if (warp < stride) {
            const uint64_t *src = lists[warp + stride];
            wide_fold_list<KPL>(R, CAP, lane, src[lane], [&](int i) -> uint64_t { return src[i]; });
        }
        __syncthreads();
        if (warp < stride) {
#pragma unroll
            for (int j = 0; j < KPL; ++j) lists[warp][j * 32 + lane] = R[j];
        }
    }
    if (warp != 0) return;
#pragma unroll
    for (int j = 0; j < KPL; ++j) {
        const int i = j * 32 + lane;
        if (i >= k) continue;
        const uint64_t key = R[j];
        const int64_t o = static_cast<int64_t>(b) * k + i;
        if (key == 0ull) {
            if (out_dist) out_dist[o] = INFINITY;
            if (out_packed) out_packed[o] = 0ull;
            out_keys[o] = -1;
            continue;
        }
        const float s = key_score(key);
        if (out_dist) out_dist[o] = l2 ? -s : 1.0f - s;
        if (out_packed) out_packed[o] = key;
        if (SHARDS) {
            out_keys[o] = shard_key_of(key, packed, shard_keys, shard_stride, b, k, cyclic_w);
        } else {
            out_keys[o] = row_keys[key_row(key)];
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
    const dim3 grid(a.B);
#define FR_MERGE(KPL, SH)                                                                               \
    merge_topk_wide_kernel<KPL, SH><<<grid, 256, 0, a.stream>>>(a.packed, a.P, a.shard_stride, a.B, a.k, a.row_keys,      \
                                                                a.shard_keys, a.l2, a.out_dist, a.out_packed, a.out_keys, \
                                                                a.only_flagged, a.limit, a.cyclic_world)
    if (a.k <= 32) {
        if (a.shards)
            merge_topk32_kernel<true><<<grid, 256, 0, a.stream>>>(a.packed, a.P, a.shard_stride, a.B, a.k, a.row_keys,
                                                                 a.shard_keys, a.l2, a.out_dist, a.out_packed, a.out_keys,
                                                                 a.only_flagged, a.limit, a.cyclic_world);
        else
            merge_topk32_kernel<false><<<grid, 256, 0, a.stream>>>(a.packed, a.P, a.shard_stride, a.B, a.k, a.row_keys,
                                                                  a.shard_keys, a.l2, a.out_dist, a.out_packed, a.out_keys,
                                                                  a.only_flagged, a.limit, a.cyclic_world);
    } else if (a.k <= 64) {
        if (a.shards) FR_MERGE(2, true); else FR_MERGE(2, false);
    } else if (a.k <= 128) {
        if (a.shards) FR_MERGE(4, true); else FR_MERGE(4, false);
    } else if (a.k <= 256 && !a.shards) {
        FR_MERGE(8, false);  // K2 second-chance lists, k' = 256
    } else {
        return cudaErrorInvalidValue;
    }
#undef FR_MERGE
    count_launch();
    return cudaGetLastError();
}

}  // namespace fr
