/*
 * nmf_oracle.c -- CPU restatement of filmyou-core's NMF / PPC clustering step.
 *
 * TEST INFRASTRUCTURE ONLY: imported by tests/ (and nothing in filmyou_core_b200/).  Plain C, fp64,
 * compiled with -ffp-contract=off (Java never fuses a*b+c).
 *
 * PINNED: tests/test_nmf_oracle_golden.py checks this file against the reference's own golden vectors
 * (T/testdata/PPCTestData.java H_one, W_one, H_ten, W_ten; T/testdata/NMFTestData.java the same;
 * T/testdata/ClusteringTestData.java H -> clustering -> clusteringCount), extracted by
 * tests/golden/make_golden_nmf.py.
 *
 * M/ = /root/reference/src/main/java/es/udc/fi/dc/irlab/.  One iteration (M/nmf/AbstractNMFDriver.java:117-138;
 * BOTH jobs read the old H and W -- hJob and wJob are constructed on (H, W, H2, W2), :118-124):
 *
 *   ComputeHJob (M/nmf/hcomputation/ComputeHJob.java:74-101)
 *     H1  X_j  = sum_i A_ij * w_i            VectorByItemHDFSMapper.java:36-58 + VectorSumReducer.java:36-47
 *     H2  C    = sum_i w_i^T w_i             CrossProductMapper.java:36-44 + MatrixSumReducer.java:35-49
 *     H3  y_j  = C h_j                       CHMapper.java:33-41
 *     H4  NMF: h_j .* x_j ./ (y_j + eps)     HComputationReducer.java:41-76
 *         PPC: d = h.y, e = h.x, h .* (x + d) ./ (y + e + eps), +-inf -> Double.MAX_VALUE
 *                                            ppc/hcomputation/PPCHComputationReducer.java:45-103
 *         The L1 renormalisation at :88-90 calls `result.normalize(1)` and DROPS the returned vector
 *         (Mahout's normalize is not in place), so the reference never normalises: H_ten's rows do not
 *         sum to 1 (golden-checked).  `apply_normalization` != 0 applies the evidently intended step.
 *   ComputeWJob (M/nmf/wcomputation/ComputeWJob.java:72-98)
 *     W1+W2  X_i = sum_j A_ij * h_j          W1Reducer.java:36-60 + VectorSumReducer
 *     W3     C   = sum_j h_j^T h_j           CrossProductMapper + MatrixSumReducer
 *     W4     y_i = w_i C                     WCMapper.java:37-48
 *     W5     w_i .* x_i ./ (y_i + eps)       WComputationMapper.java:87-116 (with the +-inf guard)
 *   eps = 1e-12                              M/nmf/MatrixComputationJob.java:41
 *
 * Summation order.  Hadoop leaves the order of a reducer's values to the shuffle, and both sum
 * reducers also run as combiners (ComputeHJob.java:129,176), so the reference's own results are
 * order-dependent at the ulp level (its tests compare at 1e-4, T/util/HadoopIntegrationTest.java:53).
 * The canonical order restated here: ratings of a row ascending by the other id (ties: input order),
 * summed in groups of `combine_len` consecutive entries (the combiner), group sums added in
 * ascending order (the reducer); cross products ascending by row id in groups of `split_rows` rows.
 * combine_len / split_rows = 0 means one plain sequential sum.
 */
#include <float.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORC_NMF_OK 0
#define ORC_NMF_E_ARG (-1)
#define ORC_NMF_E_USER_WITHOUT_RATING (-2)   /* HComputationReducer.java:52-55 "User %d has not rated any item" */
#define ORC_NMF_E_NOMEM (-6)
#define ORC_NMF_E_ITEM_WITHOUT_RATING (-10)  /* WComputationMapper.java:95-98 "Item %d has not been rated by anybody" */

static const double EPS = 1e-12;

typedef struct { int32_t row, col; float score; int64_t pos; } entry_t;

static int cmp_entry(const void* a, const void* b) {
    const entry_t* x = (const entry_t*)a; const entry_t* y = (const entry_t*)b;
    if (x->row != y->row) return x->row < y->row ? -1 : 1;
    if (x->col != y->col) return x->col < y->col ? -1 : 1;
    return x->pos < y->pos ? -1 : (x->pos > y->pos ? 1 : 0);
}

/* X[r] = sum over the entries of row r of F[col] * score, in combiner groups */
static int join_sum(const entry_t* e, int64_t m, int32_t n_rows, const double* F, int32_t k, int32_t combine_len,
                    double* X, int32_t* first_empty) {
    double* part = (double*)malloc(sizeof(double) * (size_t)k);
    if (!part) return ORC_NMF_E_NOMEM;
    int64_t p = 0;
    *first_empty = -1;
    for (int32_t r = 0; r < n_rows; r++) {
        double* x = X + (size_t)r * k;
        int64_t q = p;
        while (q < m && e[q].row == r) q++;
        if (q == p) { if (*first_empty < 0) *first_empty = r; for (int c = 0; c < k; c++) x[c] = 0.0; continue; }
        int first_group = 1;
        for (int64_t g = p; g < q;) {
            const int64_t ge = (combine_len > 0 && g + combine_len < q) ? g + combine_len : q;
            for (int64_t t = g; t < ge; t++) {
                const double* f = F + (size_t)e[t].col * k;
                const double s = (double)e[t].score;
                for (int c = 0; c < k; c++) {
                    const double v = f[c] * s;                       /* Vector.times(score) */
                    part[c] = (t == g) ? v : part[c] + v;            /* vectorAddInPlace */
                }
            }
            for (int c = 0; c < k; c++) x[c] = first_group ? part[c] : x[c] + part[c];
            first_group = 0;
            g = ge;
        }
        p = q;
    }
    free(part);
    return ORC_NMF_OK;
}

/* C = sum_r F[r]^T F[r] in groups of split_rows rows */
static int cross_sum(const double* F, int32_t n_rows, int32_t k, int32_t split_rows, double* Cm) {
    double* part = (double*)malloc(sizeof(double) * (size_t)k * k);
    if (!part) return ORC_NMF_E_NOMEM;
    int first_group = 1;
    for (int32_t g = 0; g < n_rows;) {
        const int32_t ge = (split_rows > 0 && g + split_rows < n_rows) ? g + split_rows : n_rows;
        for (int32_t r = g; r < ge; r++) {
            const double* f = F + (size_t)r * k;
            for (int a = 0; a < k; a++)
                for (int b = 0; b < k; b++) {
                    const double v = f[a] * f[b];                    /* vector.cross(vector) */
                    part[(size_t)a * k + b] = (r == g) ? v : part[(size_t)a * k + b] + v;
                }
        }
        for (size_t t = 0; t < (size_t)k * k; t++) Cm[t] = first_group ? part[t] : Cm[t] + part[t];
        first_group = 0;
        g = ge;
    }
    free(part);
    return ORC_NMF_OK;
}

static double guard_inf(double v) {                 /* Double.isInfinite(a) -> Double.MAX_VALUE (either sign) */
    return isinf(v) ? DBL_MAX : v;
}

/*
 * mode 0 = NMFDriver (ComputeHJob + ComputeWJob), 1 = PPCDriver (PPCComputeHJob + ComputeWJob).
 * Ratings: ids in [id_base, id_base + n); row r of H is user id_base + r (DataInitialization.java:82-88
 * writes keys 1..rows).  score <= 0 is dropped (VectorByItemHDFSMapper.java:40-42).  H [n_users x k],
 * W [n_items x k] row-major, updated in place over n_iter iterations.
 */
int orc_nmf_run(int32_t mode, int32_t n_users, int32_t n_items, int32_t k, int32_t id_base,
                const int32_t* user, const int32_t* item, const float* score, int64_t nnz,
                double* H, double* W, int32_t n_iter, int32_t apply_normalization, int32_t normalization_frequency,
                int32_t combine_len, int32_t split_rows, int32_t* bad_id) {
    if (n_users <= 0 || n_items <= 0 || k <= 0 || nnz < 0 || !H || !W || (mode != 0 && mode != 1)) return ORC_NMF_E_ARG;
    entry_t* by_user = (entry_t*)malloc(sizeof(entry_t) * (size_t)(nnz > 0 ? nnz : 1));
    entry_t* by_item = (entry_t*)malloc(sizeof(entry_t) * (size_t)(nnz > 0 ? nnz : 1));
    double* XH = (double*)malloc(sizeof(double) * (size_t)n_users * k);
    double* XW = (double*)malloc(sizeof(double) * (size_t)n_items * k);
    double* CW = (double*)malloc(sizeof(double) * (size_t)k * k);
    double* CH = (double*)malloc(sizeof(double) * (size_t)k * k);
    double* H2 = (double*)malloc(sizeof(double) * (size_t)n_users * k);
    double* W2 = (double*)malloc(sizeof(double) * (size_t)n_items * k);
    double* y = (double*)malloc(sizeof(double) * (size_t)k);
    int rc = ORC_NMF_OK;
    if (!by_user || !by_item || !XH || !XW || !CW || !CH || !H2 || !W2 || !y) { rc = ORC_NMF_E_NOMEM; goto done; }
    int64_t m = 0;
    for (int64_t t = 0; t < nnz; t++) {
        if (!(score[t] > 0.0f)) continue;
        const int32_t u = user[t] - id_base, i = item[t] - id_base;
        if (u < 0 || u >= n_users || i < 0 || i >= n_items) { rc = ORC_NMF_E_ARG; goto done; }
        by_user[m].row = u; by_user[m].col = i; by_user[m].score = score[t]; by_user[m].pos = t;
        by_item[m].row = i; by_item[m].col = u; by_item[m].score = score[t]; by_item[m].pos = t;
        m++;
    }
    qsort(by_user, (size_t)m, sizeof(entry_t), cmp_entry);
    qsort(by_item, (size_t)m, sizeof(entry_t), cmp_entry);

    for (int32_t it = 1; it <= n_iter; it++) {
        int32_t empty = -1;
        /* ---- ComputeHJob ---- */
        if ((rc = join_sum(by_user, m, n_users, W, k, combine_len, XH, &empty)) != ORC_NMF_OK) goto done;
        if (empty >= 0) { if (bad_id) *bad_id = empty + id_base; rc = ORC_NMF_E_USER_WITHOUT_RATING; goto done; }
        if ((rc = cross_sum(W, n_items, k, split_rows, CW)) != ORC_NMF_OK) goto done;
        /* ---- ComputeWJob inputs (old H) ---- */
        if ((rc = join_sum(by_item, m, n_items, H, k, combine_len, XW, &empty)) != ORC_NMF_OK) goto done;
        if (empty >= 0) { if (bad_id) *bad_id = empty + id_base; rc = ORC_NMF_E_ITEM_WITHOUT_RATING; goto done; }
        if ((rc = cross_sum(H, n_users, k, split_rows, CH)) != ORC_NMF_OK) goto done;

        for (int32_t j = 0; j < n_users; j++) {
            const double* h = H + (size_t)j * k; const double* x = XH + (size_t)j * k; double* o = H2 + (size_t)j * k;
            for (int r = 0; r < k; r++) {                              /* y = C.times(h): row r of C dot h */
                double s = 0.0;
                for (int c = 0; c < k; c++) s += CW[(size_t)r * k + c] * h[c];
                y[r] = s;
            }
            if (mode == 1) {
                double d = 0.0, e = 0.0;
                for (int c = 0; c < k; c++) d += h[c] * y[c];          /* vectorH.dot(vectorY) */
                for (int c = 0; c < k; c++) e += h[c] * x[c];          /* vectorH.dot(vectorX) */
                for (int c = 0; c < k; c++) {
                    const double a = guard_inf(x[c] + d), b = guard_inf(y[c] + e);
                    o[c] = h[c] * (a / (b + EPS));
                }
                if (apply_normalization && normalization_frequency != 0 && it % normalization_frequency == 0) {
                    double n1 = 0.0;
                    for (int c = 0; c < k; c++) n1 += fabs(o[c]);
                    for (int c = 0; c < k; c++) o[c] = o[c] / n1;
                }
            } else {
                for (int c = 0; c < k; c++) o[c] = h[c] * (x[c] / (y[c] + EPS));
            }
        }
        for (int32_t i = 0; i < n_items; i++) {
            const double* w = W + (size_t)i * k; const double* x = XW + (size_t)i * k; double* o = W2 + (size_t)i * k;
            for (int c = 0; c < k; c++) {                              /* y = w C : sum_r w[r] C[r][c] */
                double s = 0.0;
                for (int r = 0; r < k; r++) s += w[r] * CH[(size_t)r * k + c];
                y[c] = s;
            }
            for (int c = 0; c < k; c++) {
                const double a = guard_inf(x[c]), b = guard_inf(y[c]);
                o[c] = w[c] * (a / (b + EPS));
            }
        }
        memcpy(H, H2, sizeof(double) * (size_t)n_users * k);
        memcpy(W, W2, sizeof(double) * (size_t)n_items * k);
    }
done:
    free(by_user); free(by_item); free(XH); free(XW); free(CW); free(CH); free(H2); free(W2); free(y);
    return rc;
}

/*
 * ClusterAssignmentJob + CountClustersJob: cluster(j) = h_j.maxValueIndex()
 * (M/nmf/clustering/FindClusterMapper.java:34-42).  Mahout 0.8 AbstractVector.maxValueIndex (un-vendored
 * dependency, org.apache.mahout:mahout-math:0.8): the first strictly largest NON-ZERO element; when a zero
 * element exists and that maximum is negative (or there is no non-zero element) the first zero element.
 * cluster_size[c] = number of users assigned to c (M/nmf/clustering/CountReducer.java:35-45).
 */
int orc_cluster_assign(const double* H, int32_t n_users, int32_t k, int32_t* cluster, int32_t* cluster_size) {
    if (!H || !cluster || n_users <= 0 || k <= 0) return ORC_NMF_E_ARG;
    if (cluster_size) for (int c = 0; c < k; c++) cluster_size[c] = 0;
    for (int32_t j = 0; j < n_users; j++) {
        const double* h = H + (size_t)j * k;
        int best = -1, nz = 0, first_zero = -1;
        double mx = -INFINITY;
        for (int c = 0; c < k; c++) {
            if (h[c] == 0.0) { if (first_zero < 0) first_zero = c; continue; }
            nz++;
            if (h[c] > mx) { mx = h[c]; best = c; }
        }
        if (nz < k && mx < 0.0) best = first_zero;
        cluster[j] = best;
        if (cluster_size && best >= 0) cluster_size[best]++;
    }
    return ORC_NMF_OK;
}
