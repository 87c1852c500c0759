/*
 * fr_index.h -- C ABI of the B200-native exact vector-search backend (libfrb200.so).
 *
 * This is the drop-in boundary for ONE hot path of hawkai10/Financial-RAG: the child-chunk
 * similarity scan + top-k + cross-collection fusion that the reference runs inside the
 * third-party chromadb wheel.  Every entry point below names the reference call it replaces
 * (paths relative to the reference tree).  The reference is pure Python, so its binding is the
 * ctypes stub in INTEGRATION.md / financial_rag_b200/_lib.py.
 *
 * Conventions
 *   - extern "C", plain pointers and sizes; no torch / C++ types.
 *   - every function returns 0 on success, a negative FR_E* code otherwise; fr_last_error()
 *     returns a thread-local, NUL-terminated description of the last failure on this thread.
 *   - "host" entry points take host pointers and include the host<->device copies;
 *     "_device" entry points take device pointers on the index's device and enqueue on `stream`
 *     (a cudaStream_t passed as void*; NULL = the legacy default stream) without synchronising.
 *   - keys are int64 (the reference's Snowflake child ids, parent_child/snowflake_id.py:27-49,
 *     are int64-representable decimal strings); FR_KEY_NONE (-1) pads short result lists.
 *   - distances follow chromadb/hnswlib: cosine d = 1 - <a/|a|, b/|b|>, ip d = 1 - <a,b>,
 *     l2 d = sum (a-b)^2.  Results are sorted by ascending distance, ties by insertion order.
 *   - all entry points are thread-safe.  Mutations of one index are serialised; concurrent host searches hold the
 *     index's lock only while they are enqueued on its stream, then wait for their own event and read their own
 *     pinned staging slot (four slots per index / group).
 *   - there is NO CPU fallback: without a CUDA device every compute entry point fails.
 *   - exactness.  Results are those of an fp32-query scan of the stored rows.  The tensor-core paths SELECT candidates
 *     with bf16 query terms and certify the selection against the exact fp32 rescoring; an uncertified query is
 *     re-scanned exactly inside the same call.  The certification margin is (rigorous bound on the query rounding,
 *     |q - q_bf16| * |c|, score-aware) + (accumulation slack: 1e-5 per 384 products, x2 for two-term queries).  The
 *     slack is a statistical bound, not a worst-case one: fp32 accumulation errors of 384 products behave like
 *     sqrt(384) * 2^-24 * sum|q_i c_i| ~ 1.2e-6 (the slack is ~8 sigma); the worst case, every rounding in the same
 *     direction with truncating tensor-core accumulation, would be 384 * 2^-23 ~ 4.6e-5.
 *   - the final distances are fp32 sums of 384 products whose ORDER depends on the kernel that produced them (the
 *     streaming scan, the rescore pass after a tensor-core selection, the re-scan): the same query may return
 *     distances that differ in the last ulp (<= 2e-6 on unit vectors), and therefore a different order of rows whose
 *     exact scores tie to that precision, depending on batch size and routing.  Exact duplicates always tie exactly
 *     and come back in insertion order.
 */
#ifndef FR_INDEX_H
#define FR_INDEX_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FR_ABI_VERSION 2

typedef struct fr_index fr_index; /* opaque: one collection shard on one GPU */
typedef struct fr_group fr_group; /* opaque: one collection row-sharded over several GPUs */

/* metric vocabulary: parent_child/pgvector_child_store.py:7-26, chroma_child_store.py:34 */
enum { FR_COSINE = 0, FR_L2 = 1, FR_IP = 2 };
/* storage type of the corpus matrix in HBM */
enum { FR_BF16 = 0, FR_F32 = 1 };
/* kernel selection for fr_index_set_option("path", v): tests and bench pin a regime with it */
enum { FR_PATH_AUTO = 0, FR_PATH_STREAM = 1, FR_PATH_MMA = 2 };

/* how the shards of an fr_group bring their local top-k lists together */
enum { FR_XCHG_AUTO = 0, FR_XCHG_NCCL = 1, FR_XCHG_COPY = 2, FR_XCHG_PEER = 3 };

enum {
    FR_OK = 0,
    FR_EINVAL = -1,  /* bad argument */
    FR_ECUDA = -2,   /* CUDA runtime / driver error (message holds cudaGetErrorString) */
    FR_ENOMEM = -3,  /* host or device allocation failed */
    FR_ENODEV = -4,  /* no CUDA device / wrong architecture (needs sm_100) */
    FR_EUNSUP = -5   /* valid request this build does not implement (e.g. k > FR_MAX_K) */
};

#define FR_KEY_NONE ((int64_t)-1)
#define FR_MAX_K 128
#define FR_MAX_DIM 4096

int fr_abi_version(void);
const char *fr_last_error(void);

/* Number of kernels this library has launched in this process (all indices); bench.py reports
 * the delta over its timed region as "gpu_launches". */
int64_t fr_launch_count(void);

/* ---- collection lifecycle ---------------------------------------------------------------
 * Replaces chromadb.PersistentClient(...).get_or_create_collection(name,
 * metadata={"hnsw:space": "cosine"})            parent_child/chroma_child_store.py:32-34
 * dim: vector width (384 for bge-small / gte-small); metric: FR_COSINE|FR_L2|FR_IP;
 * dtype: FR_BF16|FR_F32; device: CUDA ordinal; reserve_rows: rows to pre-allocate (0 = grow). */
int fr_index_create(int dim, int metric, int dtype, int device, int64_t reserve_rows,
                    fr_index **out);
int fr_index_destroy(fr_index *idx);
int fr_index_reserve(fr_index *idx, int64_t rows);
int fr_index_set_option(fr_index *idx, const char *name, int64_t value);

/* Counters since creation (bench.py and tests report them): "searches", "queries", "mma_queries"
 * (queries served by the tensor-core scan), "mma_uncertified_queries" (of those, the ones whose
 * bf16-query selection could not be certified against the fp32-query order in the first pass and took
 * the second tensor-core pass) and "mma_rescanned_queries" (the ones still open after that, re-scanned
 * by the streaming kernel inside the same call).
 * Options (fr_index_set_option): "path" (FR_PATH_*), "profile", "mma_min_batch", "mma_small_max",
 * "mma_co_groups", "mma_split" (-1 auto | 0 | 1: small batches read the queries as two bf16 terms),
 * "mma_split_max", "use_graphs" (host search replays a captured CUDA graph on small collections; default 1),
 * "graph_max_bytes", "small_rows_b1" / "small_rows_b4" (FR_PATH_AUTO: collections up to this many rows send
 * batch 1 / batch <= 4 to the 3-launch streaming kernel), "mma_bound_scale_pct" (diagnostics: inflate the
 * certification bounds), "mma_wide_lists", "mma_max_lead", "mma_f32_shadow" (fp32 cosine collections of width 384:
 * tensor-core scans select on a lazily built bf16 copy of the rows, +50 % memory; default 1), "mma_score_hist" (cosine:
 * the CTAs of a tensor-core scan share a per-query score histogram for their thresholds; default 1, 0 for A/B timing),
 * "mma_debug" (bit mask, diagnostics; the bits 256 / 512 / 2048 / 4096 switch single optimisations off and leave every
 * answer unchanged, the others make results wrong), "host_debug" (diagnostics of the host path, scripts/stress_concurrent.py).
 * Further stat: "graph_replays".
 * Threading: any number of threads may call fr_index_search / fr_group_search on one object; a call holds the object's
 * lock only while it enqueues its work and then waits for its own results.  CUDA-graph replay (use_graphs) is used only
 * while no other caller of the library is waiting for results -- graph launches beside cudaEventSynchronize calls of
 * other threads crashed inside the driver (csrc/fr_host.h: graph_wait_mutex) -- otherwise the same kernels run eagerly.
 * Environment: FRB200_SEGV_TRACE=1 prints the native backtrace of a crashing thread. */
int fr_index_get_stat(fr_index *idx, const char *name, int64_t *out);

/* Replaces Collection.count()                     parent_child/chroma_child_store.py:76-80
 * count = live rows; rows = physical rows including deleted ones. */
int fr_index_count(fr_index *idx, int64_t *out_count);
int fr_index_rows(fr_index *idx, int64_t *out_rows);

/* Replaces Collection.upsert(ids, embeddings, metadatas)   chroma_child_store.py:54-59
 * vecs: n x dim row-major fp32 (host), keys: n int64 (host).  An existing key is overwritten in
 * place (keeps its insertion position); a new key is appended.  Duplicates inside one call:
 * last one wins.  Cosine collections normalise rows on the device (x * 1/(|x| + 1e-30)). */
int fr_index_upsert(fr_index *idx, const float *vecs, const int64_t *keys, int64_t n);

/* Replaces Collection.delete(ids)                 chroma_child_store.py:58, multivector_store.py:138 */
int fr_index_delete(fr_index *idx, const int64_t *keys, int64_t n, int64_t *out_deleted);

/* Bulk load from device memory (bench / sharded ingest): appends n rows without key lookup;
 * the caller guarantees the keys are new.  d_keys may be NULL: keys = first_key + i. */
int fr_index_append_device(fr_index *idx, const float *d_vecs, const int64_t *d_keys,
                           int64_t first_key, int64_t n, void *stream);

/* Read back what the index stores (tests: parity of the ingest kernel, storage-exact oracle).
 * out_vecs: n x dim fp32 (bf16 rows are widened), out_keys: n (INT64_MIN marks a deleted row). */
int fr_index_get_rows(fr_index *idx, int64_t first_row, int64_t n, float *out_vecs,
                      int64_t *out_keys);

/* ---- persistence (replaces the .chroma_children/ sqlite WAL + HNSW bin files) ------
 * The shard's rows exactly as they sit in HBM (bf16 or fp32, already normalised for cosine) and the
 * key of every physical row (INT64_MIN = deleted).  export -> a flat shard file; import appends rows
 * verbatim (no normalisation, no key lookup), so restart = mmap + H2D and a reloaded shard returns
 * bit-identical results.  out_rows / rows: n x dim x sizeof(storage type) bytes, host memory.
 * fr_index_lookup_rows: physical row of each key (-1 = absent), for incremental flushes. */
int fr_index_export_raw(fr_index *idx, int64_t first_row, int64_t n, void *out_rows,
                        int64_t *out_keys);
int fr_index_import_raw(fr_index *idx, const void *rows, const int64_t *keys, int64_t n);
int fr_index_lookup_rows(fr_index *idx, const int64_t *keys, int64_t n, int64_t *out_rows);

/* ---- the hot path -------------------------------------------------------------------------
 * Replaces Collection.query(query_embeddings=[...], n_results=k,
 *                           include=["metadatas","distances"])
 *                                  parent_child/chroma_child_store.py:63, multivector_store.py:151
 * queries: B x dim fp32; out_dist: B x k fp32 ascending; out_keys: B x k (FR_KEY_NONE padded,
 * matching pad distances are +inf).  1 <= k <= FR_MAX_K, B >= 0. */
int fr_index_search(fr_index *idx, const float *queries, int B, int k, float *out_dist,
                    int64_t *out_keys);
int fr_index_search_device(fr_index *idx, const float *d_queries, int B, int k,
                           float *d_out_dist, int64_t *d_out_keys, void *stream);

/* Row-sharded search, step 1 (per GPU): local top-k in mergeable form.  d_out_packed: B x k
 * uint64 = (order-preserving score bits << 32) | ~local_row, descending, 0 = empty slot;
 * d_out_keys: B x k int64.  Step 2 (after an all-gather of both arrays over NCCL):
 * fr_merge_shards_device merges G shards' lists, ties -> lower shard, then lower local row, so
 * the result is identical to a single-GPU search for any G.
 * d_packed: [G][B][k] uint64 with `shard_stride_bytes` between shards; same for d_keys. */
int fr_index_search_partial_device(fr_index *idx, const float *d_queries, int B, int k,
                                   uint64_t *d_out_packed, int64_t *d_out_keys, void *stream);
int fr_merge_shards_device(int device, int metric, const uint64_t *d_packed,
                           const int64_t *d_keys, int64_t shard_stride_elems, int G, int B, int k,
                           float *d_out_dist, int64_t *d_out_keys, void *stream);

/* ---- row-sharded collection (north star: "the corpus is row-sharded across the 8 GPUs of one box;
 * each GPU computes a local top-k, then an NCCL all-gather over NVLink feeds a final merge kernel") -------------
 * An fr_group is the same collection as an fr_index, spread over `world_shards` GPUs, behind the same calls:
 * it replaces the SAME reference calls (Collection.upsert / delete / query / count, chroma_child_store.py:54-63,78),
 * so get_child_vector_store(...) hands the reference's callers (rag_backend.py:632,699, retriever.py:55-56,88,
 * pipeline.py:137-143) a store whose corpus lives on all the GPUs of the box.
 *
 * Placement is cyclic: the s-th vector ever inserted (its "global row") lives on shard s % W at local row s / W; an
 * upsert of an existing key overwrites it where it is.  Results are identical to a one-GPU index for every W,
 * ties included (ascending distance, then insertion order).  A shard holds at most (2^32 - 16) / W rows.
 *
 * devices[n_local]: CUDA ordinals of the shards THIS process owns = world shards [first_shard, first_shard + n_local).
 *   One process, all GPUs (the reference's server): n_local == world_shards, first_shard 0, nccl_id NULL
 *   (ncclCommInitAll).  One process per GPU (torchrun): n_local 1, first_shard = rank, nccl_id = the 128 bytes
 *   rank 0 got from fr_nccl_unique_id and sent to everybody (ncclCommInitRank); every process then makes the same
 *   calls in the same order (SPMD), upserts included -- each keeps the rows that are its own.
 * exchange: FR_XCHG_NCCL = one grouped ncclAllGather of [packed | keys] (16 B per query and result per shard), the only
 *   mode across processes; FR_XCHG_PEER (single process) = no collective: every shard's last kernel stores its lists
 *   straight into the merging GPU's buffer through NVLink peer addressing, one event per shard orders the merge;
 *   FR_XCHG_COPY (single process) = local lists + peer copies; FR_XCHG_AUTO = PEER in one process when the devices can
 *   address each other, NCCL otherwise.
 * NCCL is bound at run time: the copy the process has already loaded (torch's bundled libnccl.so.2), else the
 * system's; fr_nccl_load(path) binds a specific one first. */
int fr_nccl_load(const char *path_or_null);
int fr_nccl_version(int *out);
int fr_nccl_unique_id(void *out, int nbytes /* >= 128 */);
int fr_group_create(int dim, int metric, int dtype, const int *devices, int n_local, int world_shards,
                    int first_shard, const void *nccl_id, int exchange, int64_t reserve_rows_per_shard,
                    fr_group **out);
int fr_group_destroy(fr_group *grp);
/* "world_shards" | "local_shards" | "first_shard" | "exchange" | "rows" | "count" | "searches" | "nccl_version" */
int fr_group_info(fr_group *grp, const char *name, int64_t *out);
/* fr_index_set_option on every local shard */
int fr_group_set_option(fr_group *grp, const char *name, int64_t value);
int fr_group_reserve(fr_group *grp, int64_t total_rows);
/* Borrowed handle of a local shard: bulk loads from device memory (fr_index_append_device with the shard's own rows
 * s, s + W, s + 2W, ... in order), per-shard options, statistics and scan profiling.  After bulk loads,
 * fr_group_adopt_rows(total) checks that every local shard holds exactly its share of `total` rows and adopts them. */
int fr_group_shard(fr_group *grp, int local_shard, fr_index **out);
int fr_group_adopt_rows(fr_group *grp, int64_t total_rows);
/* Same contracts as fr_index_count / rows / upsert / delete / get_rows / export_raw / import_raw / lookup_rows; a "row"
 * is a global row (insertion order over the whole collection), so a shard file written by a group of one size loads
 * into a group of any other size.  Row-order reads and the key map need every shard in this process. */
int fr_group_count(fr_group *grp, int64_t *out_count);
int fr_group_rows(fr_group *grp, int64_t *out_rows);
int fr_group_upsert(fr_group *grp, const float *vecs, const int64_t *keys, int64_t n);
int fr_group_delete(fr_group *grp, const int64_t *keys, int64_t n, int64_t *out_deleted);
int fr_group_get_rows(fr_group *grp, int64_t first_row, int64_t n, float *out_vecs, int64_t *out_keys);
int fr_group_export_raw(fr_group *grp, int64_t first_row, int64_t n, void *out_rows, int64_t *out_keys);
int fr_group_import_raw(fr_group *grp, const void *rows, const int64_t *keys, int64_t n);
int fr_group_lookup_rows(fr_group *grp, const int64_t *keys, int64_t n, int64_t *out_rows);
/* The hot path, same contract as fr_index_search (host buffers; the copies are inside).  With several processes the
 * one that owns shard 0 supplies `queries` (the others may pass NULL; the block is broadcast over NCCL) and every
 * process receives the result. */
int fr_group_search(fr_group *grp, const float *queries, int B, int k, float *out_dist, int64_t *out_keys);
/* Device-resident form: d_queries[j] = the B x dim query block on the device of local shard j (the same block
 * everywhere); where d_out_keys[j] / d_out_dist[j] are non-NULL the merged result is written on that device.
 * streams[j]: the stream of local shard j's device to enqueue on (streams == NULL: the group's own). */
int fr_group_search_device(fr_group *grp, const float *const *d_queries, int B, int k, float *const *d_out_dist,
                           int64_t *const *d_out_keys, void *const *streams);

/* Scan-kernel timing for the roofline line of bench.py.  After fr_index_set_option("profile", 1)
 * every search brackets its scan launches (K1 or K2, not the query normalisation or the merge)
 * with CUDA events on the stream they run on.  fr_index_profile_read waits for those events,
 * returns the summed device time, the number of scan kernel launches and of searches they
 * belong to, and resets the counters. */
int fr_index_profile_read(fr_index *idx, double *out_scan_ms, int64_t *out_scan_launches,
                          int64_t *out_searches);

/* ---- cross-collection fusion ----------------------------------------------------------------
 * Replaces the RRF loops of parent_child/retriever.py:94-107 and rag_backend.py:720-731:
 * score[key] = sum over lists, in list order, of 1.0/(k_rrf + rank) (rank from 1, fp64);
 * output sorted by score descending, ties in first-seen order, cut to k_out.
 * keys: [L][B][kp] int64 (FR_KEY_NONE entries are skipped); out_*: [B][k_out]
 * (FR_KEY_NONE / 0.0 padded). */
int fr_rrf_fuse(int device, const int64_t *keys, int L, int B, int kp, int k_rrf, int k_out,
                double *out_score, int64_t *out_keys);
int fr_rrf_fuse_device(int device, const int64_t *d_keys, int L, int B, int kp, int k_rrf,
                       int k_out, double *d_out_score, int64_t *d_out_keys, void *stream);

/* Replaces the other fusion mode of rag_backend.py:732-754 ("average of per-list min-max normalized scores"): per list
 * score = 1.0 - dist, norm = (score - min) / (max - min) (0.0 for a constant list), summed per key in list order,
 * divided by L; output sorted descending, ties in first-seen order.  fp64, bit-exact against the Python loop.
 * dist/keys: [L][B][kp] as the searches return them (key FR_KEY_NONE = empty slot); out_*: [B][k_out]. */
int fr_score_fuse(int device, const float *dist, const int64_t *keys, int L, int B, int kp, int k_out,
                  double *out_score, int64_t *out_keys);
int fr_score_fuse_device(int device, const float *d_dist, const int64_t *d_keys, int L, int B, int kp, int k_out,
                         double *d_out_score, int64_t *d_out_keys, void *stream);

/* ---- multi-vector (late interaction) aggregation ---------------------------------
 * Replaces the per-token loop of parent_child/multivector_store.py:150-176.  The T query-token vectors
 * are searched as ONE batch (fr_index_search with B = T, k = kp); this call then folds the T hit lists:
 * per token the best (1.0 - dist) of each child, summed over tokens in token order (fp64), sorted by
 * score descending with ties in first-seen order, cut to k_out.  The child ("group") of a token row is
 * key >> group_shift (the host maps "child:tok" ids to keys (child_ordinal << group_shift) | tok).
 * dist/keys: [B][T][kp] (B independent queries of T tokens; key -1 = empty slot);
 * out_score/out_group: [B][k_out], padded with 0.0 / -1. */
int fr_maxsim_aggregate(int device, const float *dist, const int64_t *keys, int B, int T, int kp,
                        int group_shift, int k_out, double *out_score, int64_t *out_group);
int fr_maxsim_aggregate_device(int device, const float *d_dist, const int64_t *d_keys, int B, int T,
                               int kp, int group_shift, int k_out, double *d_out_score,
                               int64_t *d_out_group, void *stream);

/* ---- query-side encoder ---------------------------------------------------------------------
 * Replaces the per-query BERT forward pass the reference runs on the CPU before every search:
 * SentenceTransformer(...).encode(q) at rag_backend.py:677 / parent_child/retriever.py:87, i.e. the computation
 * local_embedder.py:155-191 spells out (embeddings -> 12 transformer layers -> pooling :171-179 -> L2 normalisation :182)
 * for the two 12-layer BERT-384 encoders of the ensemble (local_models/BAAI-bge-small-en-v1.5: CLS pooling,
 * local_models/thenlper-gte-small: mean pooling; local_models/<model>/1_Pooling/config.json).  Tokenisation stays on the
 * host; the output block is laid out as fr_index_search_device / fr_group_search_device read their queries, so a
 * query batch goes from token ids to results without leaving the device.
 * Weights are handed over tensor by tensor under the key names of a Hugging Face BertModel state dict (fp32, host
 * memory); fr_encoder_finalize fails, naming what is missing, until the model is complete.
 * ids: [B][T] int32 token ids, right-padded; lens: [B] valid tokens per sequence (the attention mask);
 * pooling: 0 = [CLS] token, 1 = mean over the valid tokens; normalize: divide by max(|x|, 1e-12);
 * out: [B][hidden] fp32; hidden_or_null: optionally the last hidden state [B][T][hidden] (tests). */
typedef struct fr_encoder fr_encoder;
int fr_encoder_create(int device, int vocab_size, int hidden_size, int num_layers, int num_heads, int intermediate_size,
                      int max_position_embeddings, int type_vocab_size, float layer_norm_eps, fr_encoder **out);
int fr_encoder_destroy(fr_encoder *enc);
int fr_encoder_set_tensor(fr_encoder *enc, const char *name, const float *data, int64_t n, int *out_used);
int fr_encoder_finalize(fr_encoder *enc);
int fr_encoder_forward(fr_encoder *enc, const int32_t *ids, const int32_t *lens, int B, int T, int pooling, int normalize,
                       float *out, float *hidden_or_null);
int fr_encoder_forward_device(fr_encoder *enc, const int32_t *d_ids, const int32_t *d_lens, int B, int T, int pooling,
                              int normalize, float *d_out, float *d_hidden_or_null, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* FR_INDEX_H */
