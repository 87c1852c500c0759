// K5 rrf_fuse -- reciprocal-rank fusion across per-encoder / per-query-variant result lists.
//
// Replaces the Python loops of parent_child/retriever.py:94-107 and rag_backend.py:720-731:
//     agg[cid] = agg.get(cid, 0.0) + 1.0 / (k_rrf + rank)      (rank from 1, Python floats)
//     sorted(agg.items(), key=score, reverse=True)[:top_k]      (stable: ties keep first-seen order)
// Bit-exact by construction: fp64 IEEE division and additions in the same (list, rank) order.
// One CTA per query; candidates (L * kp <= FR_RRF_MAX_CAND) sit in shared memory; duplicate
// detection and ranking are O(n^2) compares, n <= a few hundred.  Latency-bound.
#include "fr_common.cuh"
#include "fr_kernels.h"

namespace fr {

constexpr int RRF_MAX_CAND = 2048;

__global__ void __launch_bounds__(256)
rrf_fuse_kernel(const int64_t *__restrict__ keys, int L, int B, int kp, int k_rrf, int k_out,
                double *__restrict__ out_score, int64_t *__restrict__ out_keys) {
    __shared__ int64_t ck[RRF_MAX_CAND];
    __shared__ double sc[RRF_MAX_CAND];
    __shared__ uint8_t first[RRF_MAX_CAND];
    const int b = blockIdx.x;
    const int n = L * kp;
    for (int t = threadIdx.x; t < n; t += blockDim.x) {
        const int l = t / kp, r = t - l * kp;
        ck[t] = keys[(static_cast<int64_t>(l) * B + b) * kp + r];
    }
    for (int j = threadIdx.x; j < k_out; j += blockDim.x) {
        out_keys[static_cast<int64_t>(b) * k_out + j] = -1;
        out_score[static_cast<int64_t>(b) * k_out + j] = 0.0;
    }
    __syncthreads();
    for (int t = threadIdx.x; t < n; t += blockDim.x) {
        const int64_t key = ck[t];
        bool is_first = key != -1;
        for (int u = 0; is_first && u < t; ++u) is_first = ck[u] != key;
        double s = 0.0;
        if (is_first) {
            for (int u = t; u < n; ++u)
                if (ck[u] == key) s = s + 1.0 / static_cast<double>(k_rrf + (u % kp) + 1);
        }
        first[t] = is_first ? 1 : 0;
        sc[t] = s;
    }
    __syncthreads();
    for (int t = threadIdx.x; t < n; t += blockDim.x) {
        if (!first[t]) continue;
        const double s = sc[t];
        int rank = 0;
        for (int u = 0; u < n; ++u)
            if (first[u] && (sc[u] > s || (sc[u] == s && u < t))) ++rank;
        if (rank < k_out) {
            out_keys[static_cast<int64_t>(b) * k_out + rank] = ck[t];
            out_score[static_cast<int64_t>(b) * k_out + rank] = s;
        }
    }
}

// K5b score_fuse -- the reference's other fusion mode (rag_backend.py:732-754, `fusion != "rrf"`): per list the scores
// (1.0 - dist, as the store reports them: chroma_child_store.py:72) are min-max normalised, summed per child in
// (list, rank) order and divided by the number of lists; stable sort descending, ties in first-seen order.
//     norm = (s - mn) / (mx - mn) if mx > mn else 0.0;  agg[cid] += norm;  agg[cid] /= float(len(ranked_lists))
// fp64 throughout, same operations in the same order => bit-exact against the Python loop.  A list is the run of
// non-empty entries (key != -1) at the head of its kp slots; an empty list still counts in the divisor, as the
// reference appends every search result, empty or not (rag_backend.py:701-704).
__global__ void __launch_bounds__(256)
score_fuse_kernel(const float *__restrict__ dist, const int64_t *__restrict__ keys, int L, int B, int kp, int k_out,
                  double *__restrict__ out_score, int64_t *__restrict__ out_keys) {
    __shared__ int64_t ck[RRF_MAX_CAND];
    __shared__ double sc[RRF_MAX_CAND];
    __shared__ uint8_t first[RRF_MAX_CAND];
    __shared__ double lmin[64], lmax[64];
    const int b = blockIdx.x;
    const int n = L * kp;
    for (int t = threadIdx.x; t < n; t += blockDim.x) {
        const int l = t / kp, r = t - l * kp;
        const int64_t at = (static_cast<int64_t>(l) * B + b) * kp + r;
        ck[t] = keys[at];
        sc[t] = 1.0 - static_cast<double>(dist[at]);  // score = 1.0 - float(dist)
    }
    for (int j = threadIdx.x; j < k_out; j += blockDim.x) {
        out_keys[static_cast<int64_t>(b) * k_out + j] = -1;
        out_score[static_cast<int64_t>(b) * k_out + j] = 0.0;
    }
    __syncthreads();
    for (int l = threadIdx.x; l < L; l += blockDim.x) {
        double mn = INFINITY, mx = -INFINITY;
        for (int r = 0; r < kp; ++r) {
            if (ck[l * kp + r] == -1) continue;
            const double v = sc[l * kp + r];
            mn = v < mn ? v : mn;
            mx = v > mx ? v : mx;
        }
        lmin[l] = mn;
        lmax[l] = mx;
    }
    __syncthreads();
    for (int t = threadIdx.x; t < n; t += blockDim.x) {
        const int l = t / kp;
        if (ck[t] == -1) continue;
        const double mn = lmin[l], mx = lmax[l];
        sc[t] = mx > mn ? (sc[t] - mn) / (mx - mn) : 0.0;
    }
    __syncthreads();
    double mine[RRF_MAX_CAND / 256];
#pragma unroll
    for (int i = 0; i < RRF_MAX_CAND / 256; ++i) {
        const int t = threadIdx.x + i * 256;
        double s = 0.0;
        bool is_first = false;
        if (t < n) {
            const int64_t key = ck[t];
            is_first = key != -1;
            for (int u = 0; is_first && u < t; ++u) is_first = ck[u] != key;
            if (is_first) {
                for (int u = t; u < n; ++u)
                    if (ck[u] == key) s = s + sc[u];
                s = s / static_cast<double>(L);
            }
            first[t] = is_first ? 1 : 0;
        }
        mine[i] = s;
    }
    __syncthreads();  // every thread has read the normalised scores it needs: now they may be replaced by the sums
#pragma unroll
    for (int i = 0; i < RRF_MAX_CAND / 256; ++i) {
        const int t = threadIdx.x + i * 256;
        if (t < n) sc[t] = mine[i];
    }
    __syncthreads();
    for (int t = threadIdx.x; t < n; t += blockDim.x) {
        if (!first[t]) continue;
        const double s = sc[t];
        int rank = 0;
        for (int u = 0; u < n; ++u)
            if (first[u] && (sc[u] > s || (sc[u] == s && u < t))) ++rank;
        if (rank < k_out) {
            out_keys[static_cast<int64_t>(b) * k_out + rank] = ck[t];
            out_score[static_cast<int64_t>(b) * k_out + rank] = s;
        }
    }
}

cudaError_t launch_score_fuse(const ScoreFuseArgs &a) {
    if (a.B <= 0) return cudaSuccess;
    if (a.L * a.kp > RRF_MAX_CAND || a.L > 64) return cudaErrorInvalidValue;
    score_fuse_kernel<<<a.B, 256, 0, a.stream>>>(a.dist, a.keys, a.L, a.B, a.kp, a.k_out, a.out_score, a.out_keys);
    count_launch();
    return cudaGetLastError();
}

cudaError_t launch_rrf_fuse(const RrfArgs &a) {
    if (a.B <= 0) return cudaSuccess;
    if (a.L * a.kp > RRF_MAX_CAND) return cudaErrorInvalidValue;
    rrf_fuse_kernel<<<a.B, 256, 0, a.stream>>>(a.keys, a.L, a.B, a.kp, a.k_rrf, a.k_out, a.out_score, a.out_keys);
    count_launch();
    return cudaGetLastError();
}

}  // namespace fr
