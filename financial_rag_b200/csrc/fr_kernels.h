// Host-side launch interface between api.cu and the kernel translation units.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace fr {

void count_launch();  // bumps the process-wide kernel launch counter (api.cu)

// ---- K1 scan_topk_stream ------------------------------------------------------------------
struct ScanArgs {
    const uint8_t *corpus;        // [rows(+32 pad)][dim] bf16 or fp32, row-major
    const int64_t *keys_or_null;  // non-null only when the shard holds deleted rows
    const float *queries;         // [nq_total][dim] fp32, already normalised for cosine
    int64_t n_rows;
    int dim;
    bool bf16;
    bool l2;
    int k;
    int nq_total;
    uint64_t *partials;  // [grid][nq_total][k] packed keys, written completely by the launches
    int grid;            // from scan_stream_plan_grid
    cudaStream_t stream;
};
// CTAs to launch: resident-CTA count of the least-occupant kernel variant this call will use,
// capped by the amount of work.  The same value is the number of partial lists per query.
int scan_stream_plan_grid(const ScanArgs &a, int sm_count);
cudaError_t launch_scan_stream(const ScanArgs &a);

// ---- K3 merge_topk --------------------------------------------------------------------------
struct MergeArgs {
    // mode LOCAL: `packed` = [P][B][k] partial lists of one shard (low word = ~local row);
    //             out keys come from row_keys[row].
    // mode SHARDS: `packed` = G lists per query, shard g at packed + g*shard_stride, low word
    //             ignored and replaced by ~(g*k + j); out keys from shard_keys at the same place.
    const uint64_t *packed;
    int P;
    int64_t shard_stride;  // elements between consecutive lists' base (LOCAL: B*k)
    int B;
    int k;
    bool shards;
    const int64_t *row_keys;    // LOCAL
    const int64_t *shard_keys;  // SHARDS, same layout/stride as packed
    bool l2;                    // distance = -score instead of 1 - score
    float *out_dist;            // [B][k] or null
    uint64_t *out_packed;       // [B][k] or null (mergeable form for the all-gather)
    int64_t *out_keys;          // [B][k]
    cudaStream_t stream;
};
cudaError_t launch_merge_topk(const MergeArgs &a);

// ---- K0 normalize_cast_append ----------------------------------------------------------------
struct IngestArgs {
    const float *src;            // [n][dim] fp32 (device)
    const int64_t *src_keys;     // [n] or null (then key = first_key + i)
    int64_t first_key;
    const int64_t *target_rows;  // [n] or null (then row = base_row + i); -1 = skip this vector
    int64_t base_row;
    int64_t n;
    int dim;
    bool normalize;  // cosine collections
    bool bf16;
    uint8_t *corpus;
    int64_t *keys;
    cudaStream_t stream;
};
cudaError_t launch_ingest(const IngestArgs &a);
cudaError_t launch_fill_keys(int64_t *keys, const int64_t *rows, int64_t n, int64_t value, cudaStream_t s);

// ---- K5 rrf_fuse -------------------------------------------------------------------------------
struct RrfArgs {
    const int64_t *keys;  // [L][B][kp]
    int L, B, kp, k_rrf, k_out;
    double *out_score;    // [B][k_out]
    int64_t *out_keys;    // [B][k_out]
    cudaStream_t stream;
};
cudaError_t launch_rrf_fuse(const RrfArgs &a);

}  // namespace fr
