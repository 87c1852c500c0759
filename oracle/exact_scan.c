/*
 * CPU ORACLE (test infrastructure, NOT product code) -- C restatement of the exact scan.
 *
 * Same arithmetic as oracle/exact_scan.py (which see for the reference citations:
 * parent_child/chroma_child_store.py:62-74 for the result order, chromadb/hnswlib's published
 * distance definitions for cosine / l2 / ip).  fp32 accumulation, (dist asc, row asc) order.
 * Used (a) as a second, independently written checker for the numpy oracle and (b) as the
 * multi-threaded CPU baseline that bench.py times on the GPU box's host cores.
 * PARITY STATUS: partially pinned -- see oracle/exact_scan.py's header.
 *
 * Build: make -C oracle   (gcc -O3 -march=native -fopenmp -shared -fPIC)
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

enum { FRO_COSINE = 0, FRO_L2 = 1, FRO_IP = 2 };

typedef struct {
    float d;
    int64_t row;
} cand_t;

static inline int cand_before(float d, int64_t row, const cand_t *o) {
    return d < o->d || (d == o->d && row < o->row);
}

/* insert (d,row) into a sorted (best first) list of length *len <= k */
static inline void list_insert(cand_t *lst, int *len, int k, float d, int64_t row) {
    int n = *len;
    if (n == k && !cand_before(d, row, &lst[n - 1])) return;
    int pos = n < k ? n : k - 1;
    while (pos > 0 && cand_before(d, row, &lst[pos - 1])) {
        lst[pos] = lst[pos - 1];
        --pos;
    }
    lst[pos].d = d;
    lst[pos].row = row;
    if (n < k) *len = n + 1;
}

static inline float dot_f32(const float *a, const float *b, int dim) {
    float acc = 0.0f;
#pragma omp simd reduction(+ : acc)
    for (int i = 0; i < dim; ++i) acc += a[i] * b[i];
    return acc;
}

static inline float l2sq_f32(const float *a, const float *b, int dim) {
    float acc = 0.0f;
#pragma omp simd reduction(+ : acc)
    for (int i = 0; i < dim; ++i) {
        float t = a[i] - b[i];
        acc += t * t;
    }
    return acc;
}

/* bf16 (raw uint16) -> fp32 */
static inline float bf16_to_f32(uint16_t h) {
    uint32_t u = ((uint32_t)h) << 16;
    float f;
    memcpy(&f, &u, 4);
    return f;
}

/*
 * corpus: n x dim PREPARED rows (normalised for cosine), fp32 (elem_bytes=4) or raw bf16 (2).
 * queries: B x dim PREPARED fp32.  live: optional n bytes (0 = deleted row), may be NULL.
 * out_dist/out_rows: B x k, padded with (+inf, -1).  Returns the thread count used.
 */
int fro_exact_topk(const void *corpus, int64_t n, int dim, int elem_bytes, const float *queries,
                   int B, int k, int space, const uint8_t *live, float *out_dist,
                   int64_t *out_rows, int nthreads) {
    int T = 1;
#ifdef _OPENMP
    T = nthreads > 0 ? nthreads : omp_get_max_threads();
#endif
    cand_t *lists = (cand_t *)malloc((size_t)T * B * k * sizeof(cand_t));
    int *lens = (int *)calloc((size_t)T * B, sizeof(int));
#pragma omp parallel num_threads(T)
    {
        int t = 0;
#ifdef _OPENMP
        t = omp_get_thread_num();
#endif
        cand_t *my = lists + (size_t)t * B * k;
        int *mylen = lens + (size_t)t * B;
        float *rowbuf = (float *)malloc(sizeof(float) * dim);
        int64_t lo = n * t / T, hi = n * (t + 1) / T;
        for (int64_t r = lo; r < hi; ++r) {
            if (live && !live[r]) continue;
            const float *row;
            if (elem_bytes == 4) {
                row = (const float *)corpus + r * dim;
            } else {
                const uint16_t *h = (const uint16_t *)corpus + r * dim;
                for (int i = 0; i < dim; ++i) rowbuf[i] = bf16_to_f32(h[i]);
                row = rowbuf;
            }
            for (int b = 0; b < B; ++b) {
                const float *q = queries + (size_t)b * dim;
                float d = (space == FRO_L2) ? l2sq_f32(row, q, dim) : 1.0f - dot_f32(row, q, dim);
                list_insert(my + (size_t)b * k, &mylen[b], k, d, r);
            }
        }
        free(rowbuf);
    }
    for (int b = 0; b < B; ++b) {
        cand_t *fin = (cand_t *)malloc(sizeof(cand_t) * k);
        int flen = 0;
        for (int t = 0; t < T; ++t) {
            cand_t *src = lists + ((size_t)t * B + b) * k;
            for (int j = 0; j < lens[(size_t)t * B + b]; ++j)
                list_insert(fin, &flen, k, src[j].d, src[j].row);
        }
        for (int j = 0; j < k; ++j) {
            out_dist[(size_t)b * k + j] = j < flen ? fin[j].d : INFINITY;
            out_rows[(size_t)b * k + j] = j < flen ? fin[j].row : -1;
        }
        free(fin);
    }
    free(lists);
    free(lens);
    return T;
}

/* x * 1/(||x|| + 1e-30) per row, fp32 */
void fro_normalize_rows(float *x, int64_t n, int dim) {
#pragma omp parallel for
    for (int64_t r = 0; r < n; ++r) {
        float *p = x + r * dim;
        float s = dot_f32(p, p, dim);
        float inv = 1.0f / (sqrtf(s) + 1e-30f);
        for (int i = 0; i < dim; ++i) p[i] *= inv;
    }
}
