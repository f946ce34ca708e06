/*
 * rm2_oracle.h -- CPU restatement of filmyou-core's RM2 hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * This is the parity oracle: only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference leg may load it.  The product path
 * (filmyou_core_b200/) never links, imports or calls anything in oracle/.
 *
 * The reference (100 % Java, Hadoop 1.2.1) cannot be compiled or run in this image
 * (no JVM), so this file restates its algorithm line by line in C doubles.
 * Parity is PINNED by the reference's own golden vectors: tests/test_oracle_golden.py
 * checks this oracle against all 507 (user,item,score) triples of
 * T/testdata/RMTestData.java:234-403, userSum (:408-410), totalSum (:426) and
 * itemColl (:431-464).
 *
 * Reference anchors (M/ = src/main/java/es/udc/fi/dc/irlab/):
 *   M/rm/SimpleScoreByUserHDFSMapper.java:37-40   score > 0 filter, float -> double
 *   M/rm/DoubleSumReducer.java:31-42              per-user double sum
 *   M/rm/DoubleSumAndCountReducer.java:32-45      counter += (long) sum * 100   (truncates per user)
 *   M/rm/RM2Job.java:60,95,149                    total = counter / 100
 *   M/rm/DoubleSumAndDividerReducer.java:32-46    p(i|C) = itemSum / total
 *   M/rm/AbstractRM2Reducer.java:129-233          reduce(): cluster load, P cache, user loop
 *   M/rm/AbstractRM2Reducer.java:321-371          buildRecommendations(): 3 inner loops + top-N
 *   M/rm/AbstractRM2Reducer.java:384-389          probItemGivenUser()
 *   M/util/IntDouble.java:31-34                   descending-score order (ties unspecified)
 *   M/rm/RM2HDFSReducer.java:48                   (float) score at the sink
 *
 * Orders the reference leaves to hash tables / shuffle arrival are
 * canonicalised: users ascending by id inside a cluster, items ascending by id, ties in the
 * top-N broken by ascending item id (the only total order the reference defines:
 * T/util/CassandraUtils.java:144-147, CLUSTERING ORDER BY (relevance DESC, item ASC)).
 */
#ifndef RM2_ORACLE_H
#define RM2_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum {
    ORC_MODE_LITERAL = 0,      /* row-major P[user][item], strided neighbour loop: what the JVM runs */
    ORC_MODE_LITERAL_FAST = 1, /* same arithmetic in the same order on a transposed P (bit-identical)  */
    ORC_MODE_GRAM = 2          /* G = P^T P, t = G[i][j] - P[u][i]*P[u][j]  (algebra cross-check)       */
};

enum {
    ORC_OK = 0,
    ORC_E_ARG = -1,
    ORC_E_USER_WITHOUT_RATING = -2, /* reducer would mis-parse the group, AbstractRM2Reducer.java:153-160 */
    ORC_E_DUPLICATE_RATING = -3,
    ORC_E_CLUSTER_SIZE = -4,        /* clusteringCount[c] != number of users mapped to c */
    ORC_E_UNKNOWN_USER = -5,        /* rating of a user absent from `clustering` */
    ORC_E_NOMEM = -6
};

typedef struct {
    double lambda;           /* RM2Job.LAMBDA_NAME, AbstractRM2Reducer.java:108            */
    int32_t number_of_items; /* RMRecommenderDriver.numberOfItems (global, configured) :109 */
    int32_t top_n;           /* RMRecommenderDriver.numberOfRecommendations :110            */
    int32_t filter_users;    /* RMRecommenderDriver.filterUsers :197,220-223                */
    int32_t mode;            /* ORC_MODE_*                                                  */
    int32_t threads;         /* OpenMP threads (1 = what one reduce task does)              */
    int32_t cand_stride;     /* 0/1 = every candidate; s > 1 = every s-th candidate of a user (TIMING ONLY:  */
                             /* bench.py's bounded CPU sample; time * s extrapolates to the whole user)       */
} orc_params;

typedef struct orc_result orc_result;

/* RM2-1 / RM2-2: statistics.  user_sum is indexed like `users`; item_prob by item id
 * (size max_item+1, 0 for items never rated). */
int orc_rm2_stats(const int32_t* r_user, const int32_t* r_item, const float* r_score, int64_t nnz,
                  const int32_t* users, int64_t n_users, double* user_sum,
                  int32_t max_item, double* item_sum, double* item_prob, double* total);

/* RM2-1..3 for all users (or only the users listed in only_users, when n_only > 0). */
int orc_rm2_run(const orc_params* p,
                const int32_t* r_user, const int32_t* r_item, const float* r_score, int64_t nnz,
                const int32_t* cl_user, const int32_t* cl_cluster, int64_t n_users,
                const int32_t* cluster_size, int32_t n_clusters,
                const int32_t* only_users, int64_t n_only,
                orc_result** out);

int orc_rm2_run_ext(const orc_params* p,
                    const int32_t* r_user, const int32_t* r_item, const float* r_score, int64_t nnz,
                    const int32_t* cl_user, const int32_t* cl_cluster, int64_t n_users,
                    const int32_t* cluster_size, int32_t n_clusters,
                    const int32_t* only_users, int64_t n_only,
                    const double* ext_item_prob /* may be NULL */, int32_t ext_max_item,
                    orc_result** out);

int64_t orc_result_count(const orc_result* r);
double orc_result_seconds(const orc_result* r); /* wall time of the per-user scoring loops only */
int64_t orc_result_users_scored(const orc_result* r);
/* grouped by (cluster, user id), descending score inside a user, ties by ascending item id */
void orc_result_copy(const orc_result* r, int32_t* user, int32_t* item, double* score64,
                     float* score32, int32_t* cluster);
void orc_result_free(orc_result* r);

/* Config 3: item-item co-occurrence counts C[i][j] = #users who rated both,
 * on the binarised matrix (score > 0), dense [n_items x n_items] over item ids 0..n_items-1.
 * Mahout 0.8 CooccurrenceCountSimilarity is not vendored in the reference: PARITY UNPINNED. */
int orc_cooccurrence(const int32_t* r_user, const int32_t* r_item, const float* r_score, int64_t nnz,
                     int32_t n_user_ids, int32_t n_items, int32_t* C);

#ifdef __cplusplus
}
#endif
#endif
