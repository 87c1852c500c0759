// Shared device helpers for the exact-scan kernels (sm_100a).
//
// Candidate encoding used by every kernel on the path: one uint64 "packed key"
//     (order_bits(score) << 32) | (0xFFFFFFFF - local_row)
// so a plain unsigned compare orders candidates by (score descending, row ascending) -- the
// order the reference returns (ascending distance, ties in insertion order; SURVEY.md 8c,
// observed in test_logs/query_trace_20250824_121349_f50cc515.json).  0 means "empty slot".
#pragma once

#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace fr {

constexpr unsigned FULL_MASK = 0xffffffffu;
constexpr int64_t KEY_TOMBSTONE = INT64_MIN;

__device__ __forceinline__ uint32_t order_bits(float s) {
    uint32_t u = __float_as_uint(s);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float unorder_bits(uint32_t o) {
    uint32_t u = (o & 0x80000000u) ? (o & 0x7fffffffu) : ~o;
    return __uint_as_float(u);
}
__device__ __forceinline__ uint64_t pack_key(float score, uint32_t row) {
    return (static_cast<uint64_t>(order_bits(score)) << 32) | static_cast<uint64_t>(0xffffffffu - row);
}
__device__ __forceinline__ float key_score(uint64_t key) {
    return unorder_bits(static_cast<uint32_t>(key >> 32));
}
__device__ __forceinline__ uint32_t key_row(uint64_t key) {
    return 0xffffffffu - static_cast<uint32_t>(key & 0xffffffffu);
}
// The float gate "score >= threshold_of(key)" : -inf while the list is not full.
__device__ __forceinline__ float key_threshold(uint64_t kth_key) {
    return kth_key ? key_score(kth_key) : -INFINITY;
}

// 128-bit streaming load: read-only path, do not allocate in L1 (each corpus byte is used once).
__device__ __forceinline__ uint4 ld_stream_u4(const void *p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(p));
    return r;
}

// ---- warp-wide bitonic networks on packed keys (one key per lane) -------------------------------
__device__ __forceinline__ uint64_t shfl_xor_u64(uint64_t x, int m) {
    const uint32_t lo = __shfl_xor_sync(FULL_MASK, static_cast<uint32_t>(x), m);
    const uint32_t hi = __shfl_xor_sync(FULL_MASK, static_cast<uint32_t>(x >> 32), m);
    return (static_cast<uint64_t>(hi) << 32) | lo;
}
__device__ __forceinline__ uint64_t umax64(uint64_t a, uint64_t b) { return a > b ? a : b; }
__device__ __forceinline__ uint64_t umin64(uint64_t a, uint64_t b) { return a < b ? a : b; }
// 32 keys in bitonic order -> descending (lane 0 holds the largest)
__device__ __forceinline__ uint64_t bitonic_merge32_desc(uint64_t x, int lane) {
#pragma unroll
    for (int j = 16; j > 0; j >>= 1) {
        const uint64_t y = shfl_xor_u64(x, j);
        x = ((lane & j) == 0) ? umax64(x, y) : umin64(x, y);
    }
    return x;
}
// any 32 keys -> descending
__device__ __forceinline__ uint64_t bitonic_sort32_desc(uint64_t x, int lane) {
#pragma unroll
    for (int k = 2; k <= 32; k <<= 1) {
#pragma unroll
        for (int j = k >> 1; j > 0; j >>= 1) {
            const uint64_t y = shfl_xor_u64(x, j);
            const bool desc = (lane & k) == 0;  // k = 32: every lane sorts descending
            const bool keep_max = ((lane & j) == 0) == desc;
            x = keep_max ? umax64(x, y) : umin64(x, y);
        }
    }
    return x;
}
__device__ __forceinline__ uint64_t reverse32(uint64_t x, int lane) {
    const uint32_t lo = __shfl_sync(FULL_MASK, static_cast<uint32_t>(x), 31 - lane);
    const uint32_t hi = __shfl_sync(FULL_MASK, static_cast<uint32_t>(x >> 32), 31 - lane);
    return (static_cast<uint64_t>(hi) << 32) | lo;
}

// Sort a BITONIC sequence of 32 KPL keys held striped over a warp (entry i in lane i & 31, slot i >> 5) descending:
// the strides of 32 and more pair slots of the same lane (plain register compare-exchanges), the strides below 32 are
// one shuffle stage per slot, and the slots of a stage are independent of each other (their shuffles overlap).
template <int KPL>
__device__ __forceinline__ void bitonic_merge_striped_desc(uint64_t (&L)[KPL], int lane) {
#pragma unroll
    for (int s = KPL / 2; s > 0; s >>= 1) {
#pragma unroll
        for (int j = 0; j < KPL; ++j) {
            if ((j & s) == 0) {
                const uint64_t hi = umax64(L[j], L[j + s]), lo = umin64(L[j], L[j + s]);
                L[j] = hi;
                L[j + s] = lo;
            }
        }
    }
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) {
        const bool up = (lane & s) == 0;
#pragma unroll
        for (int j = 0; j < KPL; ++j) {
            const uint64_t y = shfl_xor_u64(L[j], s);
            L[j] = up ? umax64(L[j], y) : umin64(L[j], y);
        }
    }
}

// Fold 32 keys sorted descending (one per lane, 0 = none) into a sorted list of 32 KPL entries held in registers
// (entry i in lane i & 31, slot i >> 5), keeping the best 32 KPL of the union.  Pad the block with zeros to the length of
// the list: the elementwise maxima of the list against the reversed padded block are the best 32 KPL of the union and
// form a bitonic sequence -- and only the LAST 32 entries meet a key of the block -- so the fold is one reversal, 32
// maxima and ONE merge network of log2(32 KPL) stages, of which only five shuffle.  (The first version pushed the
// block down the list 32 entries at a time, a chain of KPL x 11 dependent shuffle stages: 4 us per fold at k' = 256,
// during which the CTA pair's other seven epilogue warps and the tensor pipe wait for the accumulator to be handed
// back; profiles/r02_ncu_src_lines_scan_mma_k100_b1024_10m_before.txt.)
template <int KPL>
__device__ __forceinline__ void fold_sorted32(uint64_t (&L)[KPL], uint64_t p_sorted, int lane) {
    L[KPL - 1] = umax64(L[KPL - 1], reverse32(p_sorted, lane));
    bitonic_merge_striped_desc<KPL>(L, lane);
}

// A top-K list held by one warp in registers: entry i lives in lane (i & 31), slot (i >> 5).
// Entries are sorted descending by packed key.  KPL slots per lane => capacity 32*KPL.
template <int KPL>
struct WarpTopK {
    uint64_t e[KPL];

    __device__ __forceinline__ void clear() {
#pragma unroll
        for (int j = 0; j < KPL; ++j) e[j] = 0ull;
    }

    // Warp-uniform `key`; every lane of the warp must call.  Keeps the best `k` (<= 32*KPL).
    __device__ __forceinline__ void insert(uint64_t key, int k, int lane) {
        int pos = 0;
#pragma unroll
        for (int j = 0; j < KPL; ++j) pos += __popc(__ballot_sync(FULL_MASK, e[j] > key));
        if (pos >= k) return;  // uniform
#pragma unroll
        for (int j = KPL - 1; j >= 0; --j) {
            uint64_t up = __shfl_up_sync(FULL_MASK, e[j], 1);
            if (j > 0) {
                uint64_t carry = __shfl_sync(FULL_MASK, e[j - 1], 31);
                if (lane == 0) up = carry;
            }
            const int i = j * 32 + lane;
            e[j] = (i < pos) ? e[j] : ((i == pos) ? key : up);
        }
    }

    // entry k-1 (the current threshold), broadcast to the whole warp
    __device__ __forceinline__ uint64_t kth(int k) const {
        const int slot = (k - 1) >> 5, src = (k - 1) & 31;
        uint64_t v = 0ull;
#pragma unroll
        for (int j = 0; j < KPL; ++j) {
            uint64_t t = __shfl_sync(FULL_MASK, e[j], src);
            if (j == slot) v = t;
        }
        return v;
    }

    // store / load the whole list (32*KPL entries) to a buffer (shared or global)
    __device__ __forceinline__ void store(uint64_t *dst, int lane) const {
#pragma unroll
        for (int j = 0; j < KPL; ++j) dst[j * 32 + lane] = e[j];
    }

    // Merge a descending-sorted list `src[0..n)` (readable by all lanes) into this one.
    __device__ __forceinline__ void merge_sorted(const uint64_t *src, int n, int k, int lane) {
        uint64_t thr = kth(k);
        for (int i = 0; i < n; ++i) {
            const uint64_t key = src[i];      // same address for all lanes: broadcast
            if (key == 0ull || key <= thr) break;  // sorted: nothing further can enter
            insert(key, k, lane);
            thr = kth(k);
        }
    }
};

// Merge the per-warp lists of one CTA.  `lists` = shared memory [nwarps][nq][32*KPL];
// on return warp 0 holds the CTA-wide best k for every query.  All threads must call.
template <int KPL, int NQ>
__device__ __forceinline__ void cta_merge_lists(WarpTopK<KPL> (&tk)[NQ], uint64_t *lists, int nwarps,
                                                int warp, int lane, int k) {
    constexpr int CAP = 32 * KPL;
#pragma unroll
    for (int b = 0; b < NQ; ++b) tk[b].store(lists + (warp * NQ + b) * CAP, lane);
    for (int stride = 1; stride < nwarps; stride <<= 1) {
        __syncthreads();
        const bool active = (warp % (2 * stride) == 0) && (warp + stride < nwarps);
        if (active) {
#pragma unroll
            for (int b = 0; b < NQ; ++b)
                tk[b].merge_sorted(lists + ((warp + stride) * NQ + b) * CAP, k < CAP ? k : CAP, k, lane);
        }
        __syncthreads();
        if (active) {
#pragma unroll
            for (int b = 0; b < NQ; ++b) tk[b].store(lists + (warp * NQ + b) * CAP, lane);
        }
    }
    __syncthreads();
}

}  // namespace fr
