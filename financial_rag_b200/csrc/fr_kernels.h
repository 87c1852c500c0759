// Host-side launch interface between api.cu and the kernel translation units.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace fr {

void count_launch();  // bumps the process-wide kernel launch counter (api.cu)
void count_launches(int64_t n);  // ... by n (negative: a stream capture enqueued nothing; positive: a graph replay)

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
// Re-scan of the queries K2 could not certify: one launch, loops over fail_list[0..*fail_count),
// exits at once when the count is zero.  partials: [grid][nq_total][k], grid from
// scan_stream_fallback_grid.
bool scan_stream_fallback_serves(const ScanArgs &a);  // bf16 dot-product rows of width 384 or 768
int scan_stream_fallback_grid(const ScanArgs &a, int sm_count);
cudaError_t launch_scan_stream_fallback(const ScanArgs &a, const int *fail_count, const int *fail_list);

// ---- K2 scan_topk_mma (tcgen05) ---------------------------------------------------------------
struct MmaPlan {
    int group;       // queries per group: 128 (one CTA per SM) or 256 (CTA pairs, cta_group::2)
    int co;          // groups co-resident in a full launch (they share each corpus tile through L2)
    int lists;       // corpus streams of a full launch = partial lists per query it writes
    int tail_q0;     // first query served by the tail launch (== nq_total when there is none)
    int co_tail;     // groups in the tail launch (0 = none)
    int lists_tail;  // corpus streams of the tail launch
    int lists_max;   // partials holds [lists_max][nq_total][ksel]
};
MmaPlan scan_mma_plan(int sm_count, int64_t n_rows, int nq_total, int co_max);

// K2s score histogram of one query: 16 coarse counters (scores in [c/16, (c+1)/16)) followed by 256 fine ones
constexpr int SCORE_HIST_WORDS = 16 + 256;
struct MmaScanArgs {
    const void *corpus;         // [rows][dim] bf16
    int dim;                    // 384 for K2; K2s takes any multiple of 64 up to 1024
    const int64_t *keys_or_null;
    const void *queries_bf16;   // [nq_pad][dim * (1 + split)] bf16, zero padded (K2: nq_pad multiple of scan_mma_group())
    int split;                  // K2s only: every query row holds bf16(q) followed by bf16(q - bf16(q))
    int nq_pad;
    int64_t n_rows;
    int nq_total;
    int ksel;                   // candidates kept per query (32 or 64), >= 2k
    uint64_t *partials;         // [plan.lists_max][nq_total][ksel]; queries < plan.tail_q0 get plan.lists lists,
                                // the others plan.lists_tail
    MmaPlan plan;
    const int *nq_dev;          // optional: the launch serves *nq_dev queries (second-chance pass), 0 = exit at once
    const float *tau0;          // optional [nq_total]: fixed initial threshold per query (second-chance pass)
    uint32_t *progress;         // optional [plan.lists_max * plan.co]: tiles requested per (stream, co-resident group); lets the
                                // groups of a stream stay within max_lead tiles of each other (zeroed by the launcher)
    int max_lead;
    int dbg;                    // diagnostics only (option "mma_debug"): 1 = no corpus loads, 2 = no accumulator reads
    const float *norm2;         // K2s, l2 collections: |c|^2 per row (selection on 2 q.c - |c|^2); null otherwise
    uint32_t *tau_g;            // ksel * nq_total shared threshold slots (order_bits of a score), zeroed before the launches;
                                // K2 lays them out [ksel][nq_total], K2s [nq_total][ksel]
    uint32_t *hist;             // K2s, cosine collections, optional: [nq_total][SCORE_HIST_WORDS] score histogram shared by
                                // the CTAs (scan_mma_small.cu), zeroed before the launch
    cudaStream_t stream;
};
int scan_mma_ksel(int k, int wide = 1);  // candidates kept per query (0 = k not served by the tensor-core path); wide: 256 for k > 64
int scan_mma_group(int nq_total);  // queries per corpus pass: 128 (one CTA per SM) or 256 (CTA pairs)
struct PrepArgs {
    const float *raw;   // [nq][dim] raw fp32 queries
    int nq, nq_pad, dim;
    float *q_prep;      // [nq][dim] out: cosine-normalised fp32 queries (same arithmetic as K0)
    void *qb;           // [nq_pad][dim * (1 + split)] out: bf16 copy, zero padded; split: followed by bf16(q - bf16(q))
    float *err_bound;   // [nq] out: |q - bf16(q)|_2
    float *err_bound_split;  // [nq] out (split only): |q - bf16(q) - bf16(q - bf16(q))|_2
    float *err_alpha;        // [nq] out: |(q - bf16(q)) . q|
    float *err_alpha_split;  // [nq] out (split only): the same for the two-term residual
    int split;
    uint32_t *tau_g;    // nq * ksel threshold slots, zeroed here
    int ksel;
    uint32_t *hist;     // optional: nq * SCORE_HIST_WORDS histogram counters, zeroed here
    float bound_scale;  // >= 1: inflates the error bounds (diagnostics: forces second-chance passes; never unsafe)
    int *counters;      // n_counters ints zeroed here (failure counters of the call)
    int n_counters;
    bool normalize;     // cosine collections; inner-product collections read the raw query
    const float *cmax;  // device scalar: largest row norm (inner product), null = rows of norm <= 1.004 (cosine)
    float *inv_scale;   // [nq] out: 1 / (|q| * max row norm) -- 1 for cosine -- the scale of the absolute error slacks
    float *qnorm2;      // [nq] out or null: |q|^2 of the query as scanned (l2 certification)
    cudaStream_t stream;
};
// largest row norm of rows [0, n) folded into *out by atomicMax (a float >= 0 stored as its bits)
// norm2_out (optional, [n]): the squared norm of every row (l2 collections)
cudaError_t launch_row_norm_max(const void *rows, bool bf16, int64_t n, int dim, float *out, float *norm2_out, cudaStream_t s);
cudaError_t launch_prep_queries(const PrepArgs &a);
cudaError_t launch_scan_mma(const MmaScanArgs &a);
// K2s (scan_mma_small.cu): operands swapped for 1..64 queries, k' <= 64.  Uses plan.lists CTAs per launch;
// writes partials [plan.lists][nq_total][ksel] for queries [q0, q0 + nq).
int scan_mma_small_nq(int nq, int ksel, int dim, int split);  // padded query count (16/32/64), 0 = not served
int scan_mma_small_max_batch(int ksel, int dim, int split);   // largest batch one K2s launch serves (0 = none)
cudaError_t launch_scan_mma_small(const MmaScanArgs &a, int q0, int nq);

struct RescoreArgs {
    const uint64_t *sel;  // [B][ksel] selection lists (K3 output, packed)
    int ksel;
    const float *queries;  // [B][dim] fp32 prepared queries
    int dim;
    const uint8_t *corpus;   // the rows the collection stores: bf16, or fp32 when f32_rows
    int f32_rows;            // the scan read a bf16 copy of fp32 rows: exact scores come from the fp32 rows, and
    float extra_bound;       // the copy's rounding (<= 2^-9 sum|q_i c_i| <= 2^-9) widens the certification bound
    const float *inv_scale;  // [B] or null: see PrepArgs (absolute slacks are divided by it)
    int l2;                  // l2 collection: exact distance = sum (q - c)^2, selection scores are 2 q.c - |c|^2
    const float *qnorm2;     // l2: [B] |q|^2
    const float *cmax;       // l2: device scalar, largest row norm
    const int64_t *row_keys;
    const float *err_bound;  // [B] |e|_2, e = q - (what the scan read)
    const float *err_alpha;  // [B] |e . q|
    int split;               // the scan read two bf16 terms per query: twice the products in the accumulation slack
    int B, k;
    float *out_dist;       // [B][k] or null
    uint64_t *out_packed;  // [B][k] or null
    int64_t *out_keys;     // [B][k]
    uint8_t *flags;        // [B] 1 = not certified
    int *fail_count;
    int *fail_list;        // [B]
    unsigned long long *fail_total;  // cumulative count of uncertified queries (fr_index_get_stat)
    unsigned long long *fail_total2; // optional second counter bumped with it (failures that go straight to the re-scan)
    float *kth_exact;      // [B] out (first pass): k-th exact score, the anchor of the second-chance threshold
    const int *idx_list;   // second pass: CTA j answers query idx_list[j] from sel[j] ...
    const int *limit;      // ... for j < *limit
    const float *tau0;     // ... whose list was collected above tau0[j]
    cudaStream_t stream;
};
cudaError_t launch_rescore(const RescoreArgs &a);

// Gathers the uncertified queries of the first pass into second-chance query blocks of scan_mma_retry_max()
// queries each (scan_mma.cu): `slices` blocks hold every query that can fail, block s serves the failures
// [s * RETRY_MAX, (s + 1) * RETRY_MAX) and learns its live count from retry_n[s].
struct RetryPrepArgs {
    const float *queries;     // [B][384] fp32 prepared queries
    const float *err_bound;   // [B] |q - bf16(q)|_2 (the second-chance scan reads plain bf16 queries)
    const float *err_alpha;   // [B] |(q - bf16(q)) . q|
    float extra_bound;        // see RescoreArgs
    const float *inv_scale;   // [B] or null
    const float *kth_exact;   // [B]
    const int *fail_count;    // first-pass failures
    const int *fail_list;
    int slices;               // second-chance blocks this call enqueued (<= ceil(B / RETRY_MAX)); failures beyond them are
                              // handed to the stream re-scan: appended to fail_list2 / counted in rescanned_total
    int *fail_count2;
    int *fail_list2;
    unsigned long long *rescanned_total;
    int *host_mirror;         // optional, pinned host memory: receives the first-pass failure count of this call
    void *qb_retry;           // [slices * RETRY_MAX][384] bf16 out
    float *tau0;              // [slices * RETRY_MAX] out
    int *retry_n;             // [slices] out: live queries of each block
    uint32_t *tau_g_retry;    // [slices][ksel][RETRY_MAX], the live slices zeroed here
    int ksel;
    uint8_t *flags;
    cudaStream_t stream;
};
int scan_mma_retry_max();
int scan_mma_retry_ksel(int ksel);
cudaError_t launch_retry_prep(const RetryPrepArgs &a);

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
    const uint8_t *only_flagged;  // optional [B]: CTAs of unflagged queries exit at once (K2 fallback)
    const int *limit;             // optional: CTAs b >= *limit exit at once (K2 second-chance pass)
    int cyclic_world;             // SHARDS: 0 = shards hold contiguous row blocks (ties -> shard, position); W = global row s
                                  // lives on shard s % W at local row s / W (ties -> local_row * W + shard == global row)
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
// fp32 rows -> bf16 (RNE): the selection copy of an fp32 collection the tensor-core scans read (n_elems % 8 == 0)
cudaError_t launch_shadow_convert(const float *src, void *dst_bf16, int64_t n_elems, cudaStream_t s);

// ---- K5 rrf_fuse -------------------------------------------------------------------------------
struct RrfArgs {
    const int64_t *keys;  // [L][B][kp]
    int L, B, kp, k_rrf, k_out;
    double *out_score;    // [B][k_out]
    int64_t *out_keys;    // [B][k_out]
    cudaStream_t stream;
};
cudaError_t launch_rrf_fuse(const RrfArgs &a);
// K5b: mean of per-list min-max normalised scores (rag_backend.py:732-754)
struct ScoreFuseArgs {
    const float *dist;    // [L][B][kp] distances as the searches return them (score = 1.0 - dist)
    const int64_t *keys;  // [L][B][kp] (-1 = empty slot)
    int L, B, kp, k_out;
    double *out_score;    // [B][k_out]
    int64_t *out_keys;    // [B][k_out] (-1 padded)
    cudaStream_t stream;
};
cudaError_t launch_score_fuse(const ScoreFuseArgs &a);

// ---- K6 maxsim_aggregate -------------------------------------------------------------------------
struct MaxSimArgs {
    const float *dist;    // [B][T][kp] distances of the per-token hit lists
    const int64_t *keys;  // [B][T][kp] row keys (-1 = empty); child group = key >> group_shift
    int B, T, kp, group_shift, k_out;
    double *out_score;    // [B][k_out]
    int64_t *out_group;   // [B][k_out] (-1 padded)
    cudaStream_t stream;
};
cudaError_t launch_maxsim(const MaxSimArgs &a);

}  // namespace fr
