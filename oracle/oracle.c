/* oracle.c -- plain-C restatement of the vsearch index-scoring path.
 *
 * TEST INFRASTRUCTURE ONLY: linked/loaded by tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline leg.  Never by the product path.
 *
 * Restates (upstream jzhoubu/vsearch):
 *   Index.search           src/ir/retriever/index.py:88-94
 *       scores = q @ vector.t()      (:91)   S[b,n] = sum_j Q[b, col[n,j]] * val[n,j]
 *       scores.topk(k)               (:92)   k largest per query
 * with the contract's canonical tie rule (score desc, id asc; -0.0 == +0.0)
 * applied to the score matrix, since torch.topk's tie order is arbitrary.
 * The arithmetic of the reference lives in PyTorch/MKL (not in the reference
 * tree); this file is an independent second restatement used to cross-check
 * oracle/ref_search.py (which calls the same torch entry points as the
 * reference) against the golden vectors in tests/golden/.
 *
 * Accumulation is fp32 in CSR order (j ascending).  On dyadic-grid data every
 * order gives the same bits; on continuous data tests use a tolerance.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>

#ifdef _OPENMP
#include <omp.h>
#endif

/* (score desc, id asc) strict "a ranks before b" */
static inline int ranks_before(float sa, int64_t ia, float sb, int64_t ib) {
    if (sa > sb) return 1;
    if (sa < sb) return 0;
    return ia < ib; /* equal (incl. -0 == +0): lower id first */
}

/* min-heap on rank: root = the worst of the kept k */
typedef struct { float s; int64_t i; } ent_t;

static void sift_down(ent_t *h, int n, int p) {
    for (;;) {
        int l = 2 * p + 1, r = l + 1, w = p;
        if (l < n && ranks_before(h[w].s, h[w].i, h[l].s, h[l].i)) w = l;
        if (r < n && ranks_before(h[w].s, h[w].i, h[r].s, h[r].i)) w = r;
        if (w == p) return;
        ent_t t = h[p]; h[p] = h[w]; h[w] = t; p = w;
    }
}

static int cmp_rank(const void *a, const void *b) {
    const ent_t *x = (const ent_t *)a, *y = (const ent_t *)b;
    if (ranks_before(x->s, x->i, y->s, y->i)) return -1;
    if (ranks_before(y->s, y->i, x->s, x->i)) return 1;
    return 0;
}

static void heap_offer(ent_t *h, int *n, int k, float s, int64_t id) {
    s = s + 0.0f; /* -0.0 -> +0.0 */
    if (*n < k) {
        h[*n].s = s; h[*n].i = id; (*n)++;
        if (*n == k) for (int p = k / 2 - 1; p >= 0; --p) sift_down(h, k, p);
    } else if (ranks_before(s, id, h[0].s, h[0].i)) {
        h[0].s = s; h[0].i = id; sift_down(h, k, 0);
    }
}

static void heap_finish(ent_t *h, int n, int64_t *ids, float *sc) {
    qsort(h, (size_t)n, sizeof(ent_t), cmp_rank);
    for (int j = 0; j < n; ++j) { ids[j] = h[j].i; sc[j] = h[j].s; }
}

/* canonical top-k of a dense score matrix [B, N] (index.py:92 + tie rule) */
int oracle_topk(const float *scores, int64_t B, int64_t N, int k, int64_t *out_ids, float *out_scores) {
    if (k > N || k <= 0) return -1; /* reference: RuntimeError "selected index k out of range" */
    int rc = 0;
#pragma omp parallel for schedule(dynamic, 1)
    for (int64_t b = 0; b < B; ++b) {
        ent_t *h = (ent_t *)malloc(sizeof(ent_t) * (size_t)k);
        if (!h) { rc = -2; continue; }
        int n = 0;
        for (int64_t r = 0; r < N; ++r) heap_offer(h, &n, k, scores[b * N + r], r);
        heap_finish(h, n, out_ids + b * k, out_scores + b * k);
        free(h);
    }
    return rc;
}

/* one row of S = q . X^T for CSR X (index.py:91); val == NULL means all-ones (bag-of-token) */
static inline float csr_row_dot(const float *q, const int64_t *col, const float *val, int64_t a, int64_t e) {
    float acc = 0.0f;
    if (val) for (int64_t j = a; j < e; ++j) acc += q[col[j]] * val[j];
    else     for (int64_t j = a; j < e; ++j) acc += q[col[j]];
    return acc;
}

/* full score matrix (small inputs only) */
int oracle_csr_scores(const int64_t *crow, const int64_t *col, const float *val, int64_t N, int64_t V,
                      const float *q, int64_t B, float *out /* [B,N] */) {
    (void)V;
#pragma omp parallel for schedule(static)
    for (int64_t r = 0; r < N; ++r)
        for (int64_t b = 0; b < B; ++b)
            out[b * N + r] = csr_row_dot(q + b * V, col, val, crow[r], crow[r + 1]);
    return 0;
}

/* fused search: scores + canonical top-k without materialising [B, N].
 * Rows are split over threads; per-thread heaps are merged at the end. */
int oracle_csr_search(const int64_t *crow, const int64_t *col, const float *val, int64_t N, int64_t V,
                      const float *q, int64_t B, int k, int64_t *out_ids, float *out_scores) {
    if (k > N || k <= 0) return -1;
    int T = 1;
#ifdef _OPENMP
    T = omp_get_max_threads();
#endif
    ent_t *heaps = (ent_t *)malloc(sizeof(ent_t) * (size_t)k * (size_t)T * (size_t)B);
    int *cnt = (int *)calloc((size_t)T * (size_t)B, sizeof(int));
    if (!heaps || !cnt) { free(heaps); free(cnt); return -2; }
#pragma omp parallel
    {
        int t = 0;
#ifdef _OPENMP
        t = omp_get_thread_num();
#endif
        int64_t lo = N * t / T, hi = N * (t + 1) / T;
        for (int64_t r = lo; r < hi; ++r) {
            int64_t a = crow[r], e = crow[r + 1];
            for (int64_t b = 0; b < B; ++b) {
                float s = csr_row_dot(q + b * V, col, val, a, e);
                heap_offer(heaps + ((size_t)b * T + t) * k, &cnt[b * T + t], k, s, r);
            }
        }
    }
#pragma omp parallel for schedule(dynamic, 1)
    for (int64_t b = 0; b < B; ++b) {
        ent_t *h = (ent_t *)malloc(sizeof(ent_t) * (size_t)k);
        int n = 0;
        for (int t = 0; t < T; ++t) {
            ent_t *src = heaps + ((size_t)b * T + t) * k;
            for (int j = 0; j < cnt[b * T + t]; ++j) heap_offer(h, &n, k, src[j].s, src[j].i);
        }
        heap_finish(h, n, out_ids + b * k, out_scores + b * k);
        free(h);
    }
    free(heaps); free(cnt);
    return 0;
}

/* dense index: S = Q . X^T, X [N, D] row-major (index.py:91 with a strided vector) */
int oracle_dense_scores(const float *x, int64_t N, int64_t D, const float *q, int64_t B, float *out) {
#pragma omp parallel for schedule(static)
    for (int64_t r = 0; r < N; ++r)
        for (int64_t b = 0; b < B; ++b) {
            float acc = 0.0f;
            for (int64_t d = 0; d < D; ++d) acc += q[b * D + d] * x[r * D + d];
            out[b * N + r] = acc;
        }
    return 0;
}

int oracle_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
