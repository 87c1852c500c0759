// K6 maxsim_aggregate -- ColBERT-style late interaction over per-token nearest-neighbour lists.
//
// Replaces the Python loop of parent_child/multivector_store.py:150-176:
//     for each query token t:   local_best[child] = max over its hits of (1.0 - float(dist))
//                               child_scores[child] = child_scores.get(child, 0.0) + local_best[child]
//     sorted(child_scores.items(), key=score, reverse=True)[:top_k_children]     (stable)
// Bit-exact by construction: fp32 distance widened to fp64, "1.0 - d" and the per-child sum over tokens
// in ascending token order in fp64, ties kept in first-seen order (token, then rank inside the token).
// The T token queries themselves are ONE batched scan (B = T) instead of T single-vector queries.
// One CTA per query; T * kp <= MAXSIM_MAX_CAND candidates in shared memory; O(n^2) compares.
#include "fr_common.cuh"
#include "fr_kernels.h"

namespace fr {

constexpr int MAXSIM_MAX_CAND = 1024;

__global__ void __launch_bounds__(256)
maxsim_kernel(const float *__restrict__ dist, const int64_t *__restrict__ keys, int T, int kp, int group_shift,
              int k_out, double *__restrict__ out_score, int64_t *__restrict__ out_group) {
    __shared__ int64_t grp[MAXSIM_MAX_CAND];
    __shared__ double sc[MAXSIM_MAX_CAND];
    __shared__ double tot[MAXSIM_MAX_CAND];
    __shared__ uint8_t first[MAXSIM_MAX_CAND];
    const int b = blockIdx.x;
    const int n = T * kp;
    const int64_t base = static_cast<int64_t>(b) * n;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const int64_t key = keys[base + i];
        // key -1 pads short lists; other negative keys are legal (synthetic ids) and shift arithmetically
        grp[i] = key == -1 ? INT64_MIN : (key >> group_shift);
        sc[i] = 1.0 - static_cast<double>(dist[base + i]);
    }
    for (int j = threadIdx.x; j < k_out; j += blockDim.x) {
        out_group[static_cast<int64_t>(b) * k_out + j] = -1;
        out_score[static_cast<int64_t>(b) * k_out + j] = 0.0;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const int64_t g = grp[i];
        bool is_first = g != INT64_MIN;
        for (int u = 0; is_first && u < i; ++u) is_first = grp[u] != g;
        double s = 0.0;
        if (is_first) {
            for (int t = i / kp; t < T; ++t) {  // the child cannot occur in an earlier token
                bool any = false;
                double best = 0.0;
                for (int r = 0; r < kp; ++r) {
                    const int u = t * kp + r;
                    if (grp[u] == g && (!any || sc[u] > best)) {
                        best = sc[u];
                        any = true;
                    }
                }
                if (any) s = s + best;
            }
        }
        first[i] = is_first ? 1 : 0;
        tot[i] = s;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        if (!first[i]) continue;
        const double s = tot[i];
        int rank = 0;
        for (int u = 0; u < n; ++u)
            if (first[u] && (tot[u] > s || (tot[u] == s && u < i))) ++rank;
        if (rank < k_out) {
            out_group[static_cast<int64_t>(b) * k_out + rank] = grp[i];
            out_score[static_cast<int64_t>(b) * k_out + rank] = s;
        }
    }
}

cudaError_t launch_maxsim(const MaxSimArgs &a) {
    if (a.B <= 0) return cudaSuccess;
    if (a.T * a.kp > MAXSIM_MAX_CAND) return cudaErrorInvalidValue;
    maxsim_kernel<<<a.B, 256, 0, a.stream>>>(a.dist, a.keys, a.T, a.kp, a.group_shift, a.k_out, a.out_score,
                                            a.out_group);
    count_launch();
    return cudaGetLastError();
}

}  // namespace fr
