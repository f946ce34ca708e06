/*
 * rm2_oracle.c -- CPU restatement of filmyou-core's RM2 hot path.  TEST INFRASTRUCTURE ONLY
 * (see rm2_oracle.h for the contract, the reference anchors and how parity is pinned).
 *
 * Build (oracle/Makefile):  gcc -O2 -ffp-contract=off -fopenmp -shared -fPIC
 *   -ffp-contract=off : Java never fuses a*b+c; every operation below rounds to double
 *                       exactly as M/rm/AbstractRM2Reducer.java:344,388 does.
 */
#include "rm2_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#ifdef _OPENMP
#include <omp.h>
#endif
#if defined(__x86_64__)
#include <immintrin.h>
#endif

struct orc_result {
    int64_t count;
    int64_t users_scored;
    double seconds;
    int32_t* user;
    int32_t* item;
    double* score;
    int32_t* cluster;
};

static double now_s(void) {
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

/* ---- a rating with its dense user index, sortable by (user index, item id) ---- */
typedef struct {
    int32_t uidx; /* position of the user in the (cluster, user id)-sorted user list */
    int32_t item;
    float score;
} rating_t;

static int cmp_rating(const void* a, const void* b) {
    const rating_t* x = (const rating_t*)a;
    const rating_t* y = (const rating_t*)b;
    if (x->uidx != y->uidx) return x->uidx < y->uidx ? -1 : 1;
    if (x->item != y->item) return x->item < y->item ? -1 : 1;
    return 0;
}

typedef struct {
    int32_t id;
    int32_t cluster;
} user_t;

static int cmp_user_cluster(const void* a, const void* b) {
    const user_t* x = (const user_t*)a;
    const user_t* y = (const user_t*)b;
    if (x->cluster != y->cluster) return x->cluster < y->cluster ? -1 : 1;
    if (x->id != y->id) return x->id < y->id ? -1 : 1;
    return 0;
}

typedef struct {
    int32_t id;
    int32_t idx;
} idmap_t;

static int cmp_idmap(const void* a, const void* b) {
    const idmap_t* x = (const idmap_t*)a;
    const idmap_t* y = (const idmap_t*)b;
    return x->id < y->id ? -1 : (x->id > y->id ? 1 : 0);
}

static int32_t idmap_find(const idmap_t* m, int64_t n, int32_t id) {
    int64_t lo = 0, hi = n - 1;
    while (lo <= hi) {
        int64_t mid = (lo + hi) >> 1;
        if (m[mid].id == id) return m[mid].idx;
        if (m[mid].id < id) lo = mid + 1; else hi = mid - 1;
    }
    return -1;
}

static int cmp_i32(const void* a, const void* b) {
    int32_t x = *(const int32_t*)a, y = *(const int32_t*)b;
    return x < y ? -1 : (x > y ? 1 : 0);
}

/* ---------------------------------------------------------------------------------------------
 * GRAM mode only: G = P^T P for the columns [i0, i1) against every column j >= i0 (upper triangle, mirrored), P stored
 * transposed (column i = K contiguous doubles).  The 32 columns of an i-block stay in the core's L2 while the j columns
 * stream past once; a 4 x 2 register block of AVX2 accumulators (separate multiply and add: no FMA, like the rest of this
 * file) does the dot products.  The summation order is (v mod 4) partial sums -- GRAM is the algebra cross-check of the
 * reducer's loop, not a restatement of its order (the LITERAL modes are).
 * ------------------------------------------------------------------------------------------- */
static void gram_block_scalar(const double* P, int64_t K, int64_t I, double* G, int64_t i0, int64_t i1) {
    for (int64_t i = i0; i < i1; i++)
        for (int64_t j = i; j < I; j++) {
            const double* a = P + (size_t)i * (size_t)K;
            const double* b = P + (size_t)j * (size_t)K;
            double s0 = 0, s1 = 0, s2 = 0, s3 = 0;
            int64_t v = 0;
            for (; v + 4 <= K; v += 4) {
                s0 += a[v] * b[v]; s1 += a[v + 1] * b[v + 1];
                s2 += a[v + 2] * b[v + 2]; s3 += a[v + 3] * b[v + 3];
            }
            for (; v < K; v++) s0 += a[v] * b[v];
            const double s = (s0 + s1) + (s2 + s3);
            G[(size_t)i * (size_t)I + (size_t)j] = s;
            G[(size_t)j * (size_t)I + (size_t)i] = s;
        }
}

#if defined(__x86_64__)
__attribute__((target("avx2")))
static double hsum4(__m256d x, double tail) {
    double t[4];
    _mm256_storeu_pd(t, x);
    return ((t[0] + tail) + t[1]) + (t[2] + t[3]);
}

__attribute__((target("avx2")))
static void gram_block_avx2(const double* P, int64_t K, int64_t I, double* G, int64_t i0, int64_t i1) {
    for (int64_t j = i0; j < I; j += 2) {
        const int nj = (j + 1 < I) ? 2 : 1;
        const double* b0 = P + (size_t)j * (size_t)K;
        const double* b1 = P + (size_t)(j + nj - 1) * (size_t)K;
        for (int64_t ii = i0; ii < i1 && ii <= j + nj - 1; ii += 4) {
            const int ni = (int)((i1 - ii) < 4 ? (i1 - ii) : 4);
            const double* a[4];
            for (int q = 0; q < 4; q++) a[q] = P + (size_t)(ii + (q < ni ? q : ni - 1)) * (size_t)K;
            __m256d c00 = _mm256_setzero_pd(), c01 = c00, c10 = c00, c11 = c00, c20 = c00, c21 = c00, c30 = c00, c31 = c00;
            int64_t v = 0;
            for (; v + 4 <= K; v += 4) {
                const __m256d x0 = _mm256_loadu_pd(b0 + v), x1 = _mm256_loadu_pd(b1 + v);
                const __m256d y0 = _mm256_loadu_pd(a[0] + v), y1 = _mm256_loadu_pd(a[1] + v);
                const __m256d y2 = _mm256_loadu_pd(a[2] + v), y3 = _mm256_loadu_pd(a[3] + v);
                c00 = _mm256_add_pd(c00, _mm256_mul_pd(y0, x0)); c01 = _mm256_add_pd(c01, _mm256_mul_pd(y0, x1));
                c10 = _mm256_add_pd(c10, _mm256_mul_pd(y1, x0)); c11 = _mm256_add_pd(c11, _mm256_mul_pd(y1, x1));
                c20 = _mm256_add_pd(c20, _mm256_mul_pd(y2, x0)); c21 = _mm256_add_pd(c21, _mm256_mul_pd(y2, x1));
                c30 = _mm256_add_pd(c30, _mm256_mul_pd(y3, x0)); c31 = _mm256_add_pd(c31, _mm256_mul_pd(y3, x1));
            }
            double t[4][2] = {{0, 0}, {0, 0}, {0, 0}, {0, 0}};
            for (; v < K; v++)
                for (int q = 0; q < 4; q++) { t[q][0] += a[q][v] * b0[v]; t[q][1] += a[q][v] * b1[v]; }
            const double r[4][2] = {{hsum4(c00, t[0][0]), hsum4(c01, t[0][1])}, {hsum4(c10, t[1][0]), hsum4(c11, t[1][1])},
                                    {hsum4(c20, t[2][0]), hsum4(c21, t[2][1])}, {hsum4(c30, t[3][0]), hsum4(c31, t[3][1])}};
            for (int q = 0; q < ni; q++)
                for (int w = 0; w < nj; w++) {
                    const int64_t i = ii + q, jj = j + w;
                    if (i > jj) continue;
                    G[(size_t)i * (size_t)I + (size_t)jj] = r[q][w];
                    G[(size_t)jj * (size_t)I + (size_t)i] = r[q][w];
                }
        }
    }
}
#endif

/* ---------------------------------------------------------------------------------------------
 * LITERAL_FAST only: the neighbour sums of W = 4 G candidates against one rated item at a time.  A is the block's slice
 * of the cache, interleaved as A[v][q] (q < W); lane q accumulates  sum += P[v][i_q] * P[v][j]  for v = 0..K-1, v != u,
 * in that order, multiply and add rounded separately -- exactly the scalar loop :342-346 of each (candidate, rated item)
 * pair, W pairs side by side.  The scalar loop is bound by the latency of its one dependent add chain; W independent
 * chains fill the pipes.  ORC_LITERAL_SCALAR=1 in the environment forces the scalar loop (the tests compare both).
 * ------------------------------------------------------------------------------------------- */
#if defined(__x86_64__)
#define ORC_DOT_STEP(G)                                                                                  \
    {                                                                                                    \
        const __m256d bv = _mm256_set1_pd(b[v]);                                                         \
        const double* row = A + (size_t)v * (size_t)(4 * (G));                                           \
        _Pragma("GCC unroll 8")                                                                          \
        for (int g = 0; g < (G); g++)                                                                    \
            acc[g] = _mm256_add_pd(acc[g], _mm256_mul_pd(_mm256_loadu_pd(row + 4 * g), bv));             \
    }
#define ORC_DOT_KERNEL(G)                                                                                \
__attribute__((target("avx2")))                                                                          \
static void dot_block_##G(const double* A, const double* b, int64_t K, int64_t u, double* out) {          \
    __m256d acc[G];                                                                                      \
    _Pragma("GCC unroll 8")                                                                              \
    for (int g = 0; g < (G); g++) acc[g] = _mm256_setzero_pd();                                          \
    for (int64_t v = 0; v < u; v++) ORC_DOT_STEP(G)                                                      \
    for (int64_t v = u + 1; v < K; v++) ORC_DOT_STEP(G)                                                  \
    _Pragma("GCC unroll 8")                                                                              \
    for (int g = 0; g < (G); g++) _mm256_storeu_pd(out + 4 * g, acc[g]);                                 \
}
ORC_DOT_KERNEL(1) ORC_DOT_KERNEL(2) ORC_DOT_KERNEL(3) ORC_DOT_KERNEL(4)
ORC_DOT_KERNEL(5) ORC_DOT_KERNEL(6) ORC_DOT_KERNEL(7) ORC_DOT_KERNEL(8)
static void dot_block(int G, const double* A, const double* b, int64_t K, int64_t u, double* out) {
    switch (G) {
        case 1: dot_block_1(A, b, K, u, out); break;
        case 2: dot_block_2(A, b, K, u, out); break;
        case 3: dot_block_3(A, b, K, u, out); break;
        case 4: dot_block_4(A, b, K, u, out); break;
        case 5: dot_block_5(A, b, K, u, out); break;
        case 6: dot_block_6(A, b, K, u, out); break;
        case 7: dot_block_7(A, b, K, u, out); break;
        default: dot_block_8(A, b, K, u, out); break;
    }
}
#endif

/* ---------------------------------------------------------------------------------------------
 * Statistics: jobs RM2-1 (userSum + truncated total) and RM2-2 (p(i|C)).
 * ------------------------------------------------------------------------------------------- */
int orc_rm2_stats(const int32_t* r_user, const int32_t* r_item, const float* r_score, int64_t nnz,
                  const int32_t* users, int64_t n_users, double* user_sum,
                  int32_t max_item, double* item_sum, double* item_prob, double* total) {
    if (nnz < 0 || n_users < 0 || max_item < 0) return ORC_E_ARG;
    idmap_t* map = (idmap_t*)malloc(sizeof(idmap_t) * (size_t)(n_users > 0 ? n_users : 1));
    if (!map) return ORC_E_NOMEM;
    for (int64_t k = 0; k < n_users; k++) { map[k].id = users[k]; map[k].idx = (int32_t)k; }
    qsort(map, (size_t)n_users, sizeof(idmap_t), cmp_idmap);
    for (int64_t k = 0; k < n_users; k++) user_sum[k] = 0.0;
    for (int32_t i = 0; i <= max_item; i++) item_sum[i] = 0.0;
    for (int64_t e = 0; e < nnz; e++) {
        float s = r_score[e];
        if (!(s > 0)) continue;                          /* SimpleScoreByUserHDFSMapper.java:37-40 */
        int32_t k = idmap_find(map, n_users, r_user[e]);
        if (k < 0 || r_item[e] < 0 || r_item[e] > max_item) { free(map); return ORC_E_UNKNOWN_USER; }
        user_sum[k] += (double)s;                        /* DoubleSumReducer.java:36-38 */
        item_sum[r_item[e]] += (double)s;                /* DoubleSumAndDividerReducer.java:37-39 */
    }
    int64_t counter = 0;
    for (int64_t k = 0; k < n_users; k++)
        counter += (int64_t)user_sum[k] * 100;           /* (long) sum * OFFSET, DoubleSumAndCountReducer.java:41 */
    double t = (double)counter / 100;                    /* RM2Job.java:95,149 */
    *total = t;
    for (int32_t i = 0; i <= max_item; i++)
        item_prob[i] = item_sum[i] > 0 ? item_sum[i] / t : 0.0; /* DoubleSumAndDividerReducer.java:44 */
    free(map);
    return ORC_OK;
}

/* ---------------------------------------------------------------------------------------------
 * One scored candidate and the canonical order (score desc, item id asc).
 * ------------------------------------------------------------------------------------------- */
typedef struct {
    double score;
    int32_t item; /* original item id */
} cand_t;

static int cmp_cand(const void* a, const void* b) {
    const cand_t* x = (const cand_t*)a;
    const cand_t* y = (const cand_t*)b;
    if (x->score > y->score) return -1;                  /* IntDouble.java:31-34: Double.compare(other, this) */
    if (x->score < y->score) return 1;
    return x->item < y->item ? -1 : (x->item > y->item ? 1 : 0);
}

/* ---------------------------------------------------------------------------------------------
 * The whole job: RM2-1, RM2-2 and RM2-3 (AbstractRM2Reducer.reduce per cluster).
 * ------------------------------------------------------------------------------------------- */
int orc_rm2_run(const orc_params* p,
                const int32_t* r_user, const int32_t* r_item, const float* r_score, int64_t nnz,
                const int32_t* cl_user, const int32_t* cl_cluster, int64_t n_users,
                const int32_t* cluster_size, int32_t n_clusters,
                const int32_t* only_users, int64_t n_only,
                orc_result** out) {
    return orc_rm2_run_ext(p, r_user, r_item, r_score, nnz, cl_user, cl_cluster, n_users, cluster_size, n_clusters,
                           only_users, n_only, NULL, -1, out);
}

/* The same job with p(i|C) supplied by the caller (ext_item_prob[0..ext_max_item], NULL = computed as RM2-2 does):
 * what one reduce() call sees, where itemColl comes from the DistributedCache MapFile (AbstractRM2Reducer.java:281-303),
 * not from the group's own ratings.  Used by the neighbour-list mode of rm2_oracle.py. */
int orc_rm2_run_ext(const orc_params* p,
                    const int32_t* r_user, const int32_t* r_item, const float* r_score, int64_t nnz,
                    const int32_t* cl_user, const int32_t* cl_cluster, int64_t n_users,
                    const int32_t* cluster_size, int32_t n_clusters,
                    const int32_t* only_users, int64_t n_only,
                    const double* ext_item_prob, int32_t ext_max_item,
                    orc_result** out) {
    if (!p || !out || nnz < 0 || n_users <= 0 || n_clusters <= 0 || p->top_n < 0) return ORC_E_ARG;
    *out = NULL;
    int rc = ORC_OK;

    /* ---- users in (cluster, id) order: `users[]` of the reducer, canonicalised ---- */
    user_t* us = (user_t*)malloc(sizeof(user_t) * (size_t)n_users);
    idmap_t* map = (idmap_t*)malloc(sizeof(idmap_t) * (size_t)n_users);
    int64_t* cstart = (int64_t*)calloc((size_t)n_clusters + 1, sizeof(int64_t));
    rating_t* rt = (rating_t*)malloc(sizeof(rating_t) * (size_t)(nnz > 0 ? nnz : 1));
    int64_t* rowptr = (int64_t*)calloc((size_t)n_users + 1, sizeof(int64_t));
    double* usum = (double*)calloc((size_t)n_users, sizeof(double));
    char* wanted = NULL;
    double* isum = NULL;
    double* iprob = NULL;
    orc_result* res = NULL;
    if (!us || !map || !cstart || !rt || !rowptr || !usum) { rc = ORC_E_NOMEM; goto done; }

    for (int64_t k = 0; k < n_users; k++) {
        us[k].id = cl_user[k];
        us[k].cluster = cl_cluster[k];
        if (cl_cluster[k] < 0 || cl_cluster[k] >= n_clusters) { rc = ORC_E_ARG; goto done; }
    }
    qsort(us, (size_t)n_users, sizeof(user_t), cmp_user_cluster);
    for (int64_t k = 0; k < n_users; k++) {
        map[k].id = us[k].id; map[k].idx = (int32_t)k;
        cstart[us[k].cluster + 1]++;
    }
    for (int32_t c = 0; c < n_clusters; c++) {
        if (cstart[c + 1] != cluster_size[c]) { rc = ORC_E_CLUSTER_SIZE; goto done; }
        cstart[c + 1] += cstart[c];
    }
    qsort(map, (size_t)n_users, sizeof(idmap_t), cmp_idmap);
    for (int64_t k = 1; k < n_users; k++)
        if (map[k].id == map[k - 1].id) { rc = ORC_E_ARG; goto done; }

    /* ---- positive ratings only, by (user index, item) ---- */
    int64_t m = 0;
    int32_t max_item = 0;
    for (int64_t e = 0; e < nnz; e++) {
        if (!(r_score[e] > 0)) continue;                 /* ScoreByClusterHDFSMapper.java:39-40 */
        int32_t k = idmap_find(map, n_users, r_user[e]);
        if (k < 0) { rc = ORC_E_UNKNOWN_USER; goto done; }
        if (r_item[e] < 0) { rc = ORC_E_ARG; goto done; }
        rt[m].uidx = k; rt[m].item = r_item[e]; rt[m].score = r_score[e];
        if (r_item[e] > max_item) max_item = r_item[e];
        m++;
    }
    qsort(rt, (size_t)m, sizeof(rating_t), cmp_rating);
    for (int64_t e = 1; e < m; e++)
        if (rt[e].uidx == rt[e - 1].uidx && rt[e].item == rt[e - 1].item) { rc = ORC_E_DUPLICATE_RATING; goto done; }
    for (int64_t e = 0; e < m; e++) rowptr[rt[e].uidx + 1]++;
    for (int64_t k = 0; k < n_users; k++) {
        if (rowptr[k + 1] == 0) { rc = ORC_E_USER_WITHOUT_RATING; goto done; }
        rowptr[k + 1] += rowptr[k];
    }

    /* ---- RM2-1: user sums (ascending item order) and the truncated total ---- */
    int64_t counter = 0;
    for (int64_t k = 0; k < n_users; k++) {
        double s = 0;
        for (int64_t e = rowptr[k]; e < rowptr[k + 1]; e++) s += (double)rt[e].score;
        usum[k] = s;
        counter += (int64_t)s * 100;                     /* DoubleSumAndCountReducer.java:41 */
    }
    const double total = (double)counter / 100;          /* RM2Job.java:95 */

    /* ---- RM2-2: p(i|C), item sums accumulated in ascending (cluster,user) order ---- */
    isum = (double*)calloc((size_t)max_item + 1, sizeof(double));
    iprob = (double*)calloc((size_t)max_item + 1, sizeof(double));
    if (!isum || !iprob) { rc = ORC_E_NOMEM; goto done; }
    for (int64_t e = 0; e < m; e++) isum[rt[e].item] += (double)rt[e].score;
    for (int32_t i = 0; i <= max_item; i++) iprob[i] = isum[i] / total; /* DoubleSumAndDividerReducer.java:44 */
    if (ext_item_prob) {                                 /* itemColl handed in, as the reducer reads it from the cache file */
        if (ext_max_item < max_item) { rc = ORC_E_ARG; goto done; }   /* "p(i|C) not found", :294-298 */
        for (int32_t i = 0; i <= max_item; i++) iprob[i] = ext_item_prob[i];
    }

    /* ---- which users are scored ---- */
    wanted = (char*)malloc((size_t)n_users);
    if (!wanted) { rc = ORC_E_NOMEM; goto done; }
    if (n_only > 0) {
        memset(wanted, 0, (size_t)n_users);
        for (int64_t k = 0; k < n_only; k++) {
            int32_t idx = idmap_find(map, n_users, only_users[k]);
            if (idx < 0) { rc = ORC_E_UNKNOWN_USER; goto done; }
            wanted[idx] = 1;
        }
    } else {
        memset(wanted, 1, (size_t)n_users);
    }

    res = (orc_result*)calloc(1, sizeof(orc_result));
    if (!res) { rc = ORC_E_NOMEM; goto done; }
    int64_t cap = 0;
    {
        int64_t nw = 0;
        for (int64_t k = 0; k < n_users; k++) nw += wanted[k];
        /* upper bound per user: min(top_n, max_item+1) */
        int64_t per = (int64_t)max_item + 1;
        if (p->top_n < per) per = p->top_n;
        cap = nw * per + 1;
    }
    res->user = (int32_t*)malloc(sizeof(int32_t) * (size_t)cap);
    res->item = (int32_t*)malloc(sizeof(int32_t) * (size_t)cap);
    res->score = (double*)malloc(sizeof(double) * (size_t)cap);
    res->cluster = (int32_t*)malloc(sizeof(int32_t) * (size_t)cap);
    if (!res->user || !res->item || !res->score || !res->cluster) { rc = ORC_E_NOMEM; goto done; }

    const double lambda = p->lambda;
    const double log_items = log((double)p->number_of_items);   /* AbstractRM2Reducer.java:328 */
    int threads = p->threads > 0 ? p->threads : 1;
#ifdef _OPENMP
    omp_set_num_threads(threads);
#else
    (void)threads;
#endif

    /* ---- RM2-3: one "reduce" per cluster ---- */
    for (int32_t c = 0; c < n_clusters && rc == ORC_OK; c++) {
        const int64_t u0 = cstart[c], u1 = cstart[c + 1];
        const int64_t K = u1 - u0;                       /* numberOfUsersInCluster :143 */
        if (K == 0) continue;
        int any = 0;
        for (int64_t k = u0; k < u1; k++) any |= wanted[k];
        if (!any) continue;

        /* items[] = items with >= 1 positive rating from the cluster (:164-174, :238-243), ascending */
        const int64_t e0 = rowptr[u0], e1 = rowptr[u1];
        int32_t* items = (int32_t*)malloc(sizeof(int32_t) * (size_t)(e1 - e0 + 1));
        int32_t* loc = (int32_t*)malloc(sizeof(int32_t) * ((size_t)max_item + 1));
        if (!items || !loc) { free(items); free(loc); rc = ORC_E_NOMEM; break; }
        for (int64_t e = e0; e < e1; e++) items[e - e0] = rt[e].item;
        qsort(items, (size_t)(e1 - e0), sizeof(int32_t), cmp_i32);
        int64_t I = 0;
        for (int64_t e = 0; e < e1 - e0; e++)
            if (e == 0 || items[e] != items[e - 1]) items[I++] = items[e];
        for (int64_t i = 0; i < I; i++) loc[items[i]] = (int32_t)i;

        /* cache[u][i] = probItemGivenUser(i,u) (:184-190, :384-389) */
        const int transposed = (p->mode != ORC_MODE_LITERAL);
        double* P = (double*)malloc(sizeof(double) * (size_t)K * (size_t)I);
        double* G = NULL;
        if (!P) { free(items); free(loc); rc = ORC_E_NOMEM; break; }
#define PAT(v, i) (transposed ? P[(size_t)(i) * (size_t)K + (size_t)(v)] : P[(size_t)(v) * (size_t)I + (size_t)(i)])
        for (int64_t v = 0; v < K; v++) {
            const double sum = usum[u0 + v];
            for (int64_t i = 0; i < I; i++) {
                const double val = (1 - lambda) * (0.0 / sum) + lambda * iprob[items[i]];
                if (transposed) P[(size_t)i * (size_t)K + (size_t)v] = val; else P[(size_t)v * (size_t)I + (size_t)i] = val;
            }
            for (int64_t e = rowptr[u0 + v]; e < rowptr[u0 + v + 1]; e++) {
                const int64_t i = loc[rt[e].item];
                const double rating = (double)rt[e].score;                      /* :173 */
                const double val = (1 - lambda) * (rating / sum) + lambda * iprob[items[i]]; /* :388 */
                if (transposed) P[(size_t)i * (size_t)K + (size_t)v] = val; else P[(size_t)v * (size_t)I + (size_t)i] = val;
            }
        }
        if (p->mode == ORC_MODE_GRAM) {
            G = (double*)malloc(sizeof(double) * (size_t)I * (size_t)I);
            if (!G) { free(P); free(items); free(loc); rc = ORC_E_NOMEM; break; }
            {
                int use_avx2 = 0;
#if defined(__x86_64__)
                use_avx2 = __builtin_cpu_supports("avx2");
#endif
                const int64_t IB = 32;                         /* columns of an i-block: 32 x K doubles stay in L2 */
                const int64_t n_blocks = (I + IB - 1) / IB;
#pragma omp parallel for schedule(dynamic, 1)
                for (int64_t bk = 0; bk < n_blocks; bk++) {
                    const int64_t i0 = bk * IB, i1 = (i0 + IB < I) ? i0 + IB : I;
#if defined(__x86_64__)
                    if (use_avx2) { gram_block_avx2(P, K, I, G, i0, i1); continue; }
#endif
                    gram_block_scalar(P, K, I, G, i0, i1);
                }
                (void)use_avx2;
            }
        }

        /* per-user result slots so that the parallel loop is deterministic */
        int64_t* slot = (int64_t*)malloc(sizeof(int64_t) * (size_t)(K + 1));
        int64_t base = res->count, nslots = 0;
        for (int64_t v = 0; v < K; v++) {
            slot[v] = -1;
            const int64_t n = rowptr[u0 + v + 1] - rowptr[u0 + v];
            const int64_t cu = I - n;
            if (!wanted[u0 + v]) continue;
            if (cu == 0) continue;                        /* :210-213 */
            if (us[u0 + v].id < p->filter_users) continue; /* :220-223 */
            slot[v] = base + nslots;
            nslots += (cu < p->top_n ? cu : p->top_n);
        }
        const double log_K = log((double)K);              /* :329 */
        const double t_begin = now_s();
        int oom = 0;
        /* Users are independent (one reduce task scores them one after the other, :202-226).  To keep every host
         * thread busy whatever the mix of light and heavy users, the (user, block of 64 candidates) pairs are the
         * parallel tasks; the arithmetic of one (user, candidate) pair -- the loops :337-348 -- is untouched.
         * cand_stride > 1 (timing only, bench.py) scores every stride-th candidate of a user: the cost of a
         * candidate does not depend on which one it is, so time * stride extrapolates to the whole user. */
        const int64_t stride = p->cand_stride > 1 ? p->cand_stride : 1;
        typedef struct { cand_t* prefs; int32_t* rj; int32_t* cand_i; int64_t np; int64_t task0; int n; } ujob_t;
        ujob_t* job = (ujob_t*)calloc((size_t)K, sizeof(ujob_t));
        int64_t n_tasks = 0;
        const int64_t CH = 64;
        if (!job) { oom = 1; }
        for (int64_t u = 0; u < K && !oom; u++) {
            if (slot[u] < 0) continue;
            const int64_t r0 = rowptr[u0 + u], r1 = rowptr[u0 + u + 1];
            const int n = (int)(r1 - r0);
            const int64_t cu = I - n;
            char* rated = (char*)calloc((size_t)I, 1);
            job[u].n = n;
            job[u].rj = (int32_t*)malloc(sizeof(int32_t) * (size_t)n);
            job[u].cand_i = (int32_t*)malloc(sizeof(int32_t) * (size_t)(cu > 0 ? cu : 1));
            job[u].prefs = (cand_t*)malloc(sizeof(cand_t) * (size_t)(cu > 0 ? cu : 1));
            if (!rated || !job[u].rj || !job[u].cand_i || !job[u].prefs) { oom = 1; free(rated); break; }
            for (int k = 0; k < n; k++) { job[u].rj[k] = loc[rt[r0 + k].item]; rated[job[u].rj[k]] = 1; }
            int64_t np = 0, seen = 0;
            for (int64_t i = 0; i < I; i++)                                     /* unratedItems :206-208 */
                if (!rated[i]) { if (seen % stride == 0) job[u].cand_i[np++] = (int32_t)i; seen++; }
            job[u].np = np;
            job[u].task0 = n_tasks;
            n_tasks += (np + CH - 1) / CH;
            free(rated);
        }
        int64_t* task_user = NULL;
        if (!oom) {
            task_user = (int64_t*)malloc(sizeof(int64_t) * (size_t)(n_tasks > 0 ? n_tasks : 1));
            if (!task_user) oom = 1;
            else for (int64_t u = 0; u < K; u++)
                if (slot[u] >= 0) for (int64_t t = job[u].task0, e = job[u].task0 + (job[u].np + CH - 1) / CH; t < e; t++) task_user[t] = u;
        }
        if (!oom) {
            /* candidates per interleaved block: the block (K x W doubles) should stay in a core's L2 */
            int block_w = 0;
#if defined(__x86_64__)
            if (p->mode == ORC_MODE_LITERAL_FAST && __builtin_cpu_supports("avx2") && !getenv("ORC_LITERAL_SCALAR")) {
                int64_t w = ((int64_t)768 * 1024 / (8 * K)) & ~(int64_t)3;
                block_w = (int)(w < 4 ? 4 : (w > 32 ? 32 : w));
            }
#endif
#pragma omp parallel
            {
            double* blockA = block_w ? (double*)malloc(sizeof(double) * (size_t)K * (size_t)block_w) : NULL;
#pragma omp for schedule(dynamic, 1)
            for (int64_t t = 0; t < n_tasks; t++) {
                const int64_t u = task_user[t];
                const ujob_t* jb = &job[u];
                const int n = jb->n;
                const int32_t* rj = jb->rj;
                const double pvpi = (n - 1) * log_items - n * log_K;           /* :328-329 */
                const int64_t q0 = (t - jb->task0) * CH, q1 = q0 + CH < jb->np ? q0 + CH : jb->np;
#if defined(__x86_64__)
                if (blockA) {
                    double acc[64], sums[32];
                    for (int64_t q = q0; q < q1; q++) acc[q - q0] = 0.0;
                    for (int64_t qb = q0; qb < q1; qb += block_w) {
                        const int nq = (int)(q1 - qb < block_w ? q1 - qb : block_w);
                        const int Gn = (nq + 3) / 4, W = 4 * Gn;
                        for (int q = 0; q < W; q++) {                           /* pad lanes repeat the last candidate */
                            const double* a = P + (size_t)jb->cand_i[qb + (q < nq ? q : nq - 1)] * (size_t)K;
                            for (int64_t v = 0; v < K; v++) blockA[(size_t)v * (size_t)W + (size_t)q] = a[v];
                        }
                        for (int k = 0; k < n; k++) {                           /* :337 */
                            dot_block(Gn, blockA, P + (size_t)rj[k] * (size_t)K, K, u, sums);   /* :339-346 */
                            for (int q = 0; q < nq; q++) acc[qb - q0 + q] += log(sums[q]);       /* :348 */
                        }
                    }
                    for (int64_t q = q0; q < q1; q++) {
                        jb->prefs[q].score = acc[q - q0] + pvpi;                /* :352 */
                        jb->prefs[q].item = items[jb->cand_i[q]];
                    }
                    continue;
                }
#endif
                if (p->mode == ORC_MODE_GRAM) {
                    /* The arithmetic of every (candidate, rated item) pair and the k-ascending accumulation per candidate are
                     * those of the loop below, with k outermost: G is symmetric (both halves were stored from the same
                     * value), so row j is read contiguously over the block's candidates instead of one cache line per term. */
                    double acc[64], pui[64];
                    for (int64_t q = q0; q < q1; q++) { acc[q - q0] = 0.0; pui[q - q0] = PAT(u, jb->cand_i[q]); }
                    for (int k = 0; k < n; k++) {
                        const int64_t j = rj[k];
                        const double* Gj = G + (size_t)j * (size_t)I;
                        const double puj = PAT(u, j);
                        for (int64_t q = q0; q < q1; q++) {
                            const double sum = Gj[jb->cand_i[q]] - pui[q - q0] * puj;
                            acc[q - q0] += log(sum);                           /* :348 */
                        }
                    }
                    for (int64_t q = q0; q < q1; q++) {
                        jb->prefs[q].score = acc[q - q0] + pvpi;                /* :352 */
                        jb->prefs[q].item = items[jb->cand_i[q]];
                    }
                    continue;
                }
                for (int64_t q = q0; q < q1; q++) {                             /* :332 */
                    const int64_t i = jb->cand_i[q];
                    double logResult = 0.0;                                    /* :334 */
                    for (int k = 0; k < n; k++) {                               /* :337 */
                        const int64_t j = rj[k];
                        double sum = 0.0;                                      /* :339 */
                        if (p->mode == ORC_MODE_GRAM) {
                            sum = G[(size_t)i * (size_t)I + (size_t)j] - PAT(u, i) * PAT(u, j);
                        } else if (transposed) {
                            const double* a = P + (size_t)i * (size_t)K;
                            const double* b = P + (size_t)j * (size_t)K;
                            for (int64_t v = 0; v < u; v++) sum += a[v] * b[v];
                            for (int64_t v = u + 1; v < K; v++) sum += a[v] * b[v];
                        } else {
                            for (int64_t v = 0; v < K; v++) {                  /* :342-346 */
                                if (v == u) continue;                          /* neighbours.remove(userID) :216 */
                                sum += P[(size_t)v * (size_t)I + (size_t)i] * P[(size_t)v * (size_t)I + (size_t)j];
                            }
                        }
                        logResult += log(sum);                                 /* :348 */
                    }
                    logResult += pvpi;                                         /* :352 */
                    jb->prefs[q].score = logResult;
                    jb->prefs[q].item = items[i];
                }
            }
            free(blockA);
            }
#pragma omp parallel for schedule(dynamic, 1)
            for (int64_t u = 0; u < K; u++) {
                if (slot[u] < 0) continue;
                qsort(job[u].prefs, (size_t)job[u].np, sizeof(cand_t), cmp_cand);  /* PriorityQueue poll order :358-369 */
                int64_t iterations = job[u].np < p->top_n ? job[u].np : p->top_n;  /* :360 */
                const int64_t room = (I - job[u].n) < p->top_n ? (I - job[u].n) : p->top_n;
                for (int64_t k = 0; k < room; k++) {                            /* :361-369 */
                    const int64_t kk = k < iterations ? k : iterations - 1;     /* stride > 1: pad (timing mode only) */
                    res->user[slot[u] + k] = us[u0 + u].id;
                    res->item[slot[u] + k] = job[u].prefs[kk].item;
                    res->score[slot[u] + k] = job[u].prefs[kk].score;
                    res->cluster[slot[u] + k] = c;
                }
            }
        }
        if (job) for (int64_t u = 0; u < K; u++) { free(job[u].prefs); free(job[u].rj); free(job[u].cand_i); }
        free(job); free(task_user);
        res->seconds += now_s() - t_begin;
        for (int64_t v = 0; v < K; v++) if (slot[v] >= 0) res->users_scored++;
        res->count += nslots;
        if (oom) rc = ORC_E_NOMEM;
        free(slot); free(G); free(P); free(items); free(loc);
#undef PAT
    }

done:
    free(us); free(map); free(cstart); free(rt); free(rowptr); free(usum);
    free(wanted); free(isum); free(iprob);
    if (rc != ORC_OK) { orc_result_free(res); return rc; }
    *out = res;
    return ORC_OK;
}

int64_t orc_result_count(const orc_result* r) { return r ? r->count : 0; }
double orc_result_seconds(const orc_result* r) { return r ? r->seconds : 0.0; }
int64_t orc_result_users_scored(const orc_result* r) { return r ? r->users_scored : 0; }

void orc_result_copy(const orc_result* r, int32_t* user, int32_t* item, double* score64,
                     float* score32, int32_t* cluster) {
    if (!r) return;
    for (int64_t k = 0; k < r->count; k++) {
        if (user) user[k] = r->user[k];
        if (item) item[k] = r->item[k];
        if (score64) score64[k] = r->score[k];
        if (score32) score32[k] = (float)r->score[k];     /* RM2HDFSReducer.java:48 */
        if (cluster) cluster[k] = r->cluster[k];
    }
}

void orc_result_free(orc_result* r) {
    if (!r) return;
    free(r->user); free(r->item); free(r->score); free(r->cluster);
    free(r);
}

/* ---------------------------------------------------------------------------------------------
 * Config 3: item-item co-occurrence counts on the binarised matrix.  PARITY UNPINNED (Mahout 0.8
 * RowSimilarityJob/CooccurrenceCountSimilarity, called at
 * M/baselinerecommender/BaselineRecommenderJob.java:241-253, is not in /root/reference);
 * semantics restated from Mahout's published definition: C[i][j] = sum_u B[u][i]*B[u][j].
 * ------------------------------------------------------------------------------------------- */
typedef struct { int32_t user, item; } ui_t;
static int cmp_ui(const void* a, const void* b) {
    const ui_t* x = (const ui_t*)a; const ui_t* y = (const ui_t*)b;
    if (x->user != y->user) return x->user < y->user ? -1 : 1;
    return x->item < y->item ? -1 : (x->item > y->item ? 1 : 0);
}

int orc_cooccurrence(const int32_t* r_user, const int32_t* r_item, const float* r_score, int64_t nnz,
                     int32_t n_user_ids, int32_t n_items, int32_t* C) {
    if (nnz < 0 || n_items <= 0 || n_user_ids <= 0) return ORC_E_ARG;
    ui_t* a = (ui_t*)malloc(sizeof(ui_t) * (size_t)(nnz > 0 ? nnz : 1));
    if (!a) return ORC_E_NOMEM;
    int64_t m = 0;
    for (int64_t e = 0; e < nnz; e++) {
        if (!(r_score[e] > 0)) continue;
        if (r_item[e] < 0 || r_item[e] >= n_items || r_user[e] < 0 || r_user[e] >= n_user_ids) { free(a); return ORC_E_ARG; }
        a[m].user = r_user[e]; a[m].item = r_item[e]; m++;
    }
    qsort(a, (size_t)m, sizeof(ui_t), cmp_ui);
    memset(C, 0, sizeof(int32_t) * (size_t)n_items * (size_t)n_items);
    int64_t s = 0;
    while (s < m) {
        int64_t t = s;
        while (t < m && a[t].user == a[s].user) t++;
        for (int64_t x = s; x < t; x++) {
            if (x > s && a[x].item == a[x - 1].item) continue;     /* binarised: duplicates count once */
            for (int64_t y = s; y < t; y++) {
                if (y > s && a[y].item == a[y - 1].item) continue;
                C[(size_t)a[x].item * (size_t)n_items + (size_t)a[y].item] += 1;
            }
        }
        s = t;
    }
    free(a);
    return ORC_OK;
}
