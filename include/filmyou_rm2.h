/*
 * filmyou_rm2.h -- C ABI of the B200-native RM2 engine (libfilmyou_rm2.so).
 *
 * Drop-in boundary for filmyou-core's job chain RM2-1..3.  The reference exposes no FFI of its
 * own (it is 100 % Java), so each entry point below names the reference interface it replaces.
 * M/ = /root/reference/src/main/java/es/udc/fi/dc/irlab/.
 *
 *   coarse seam: RM2Job.run                         M/rm/RM2Job.java:76-100
 *       fy_rm2_set_ratings    <- the ratings SequenceFile<IntPairWritable(user,item), FloatWritable>
 *                                read by the three mappers (M/rm/RM2Job.java:118-133,170-185,220-235)
 *       fy_rm2_set_clustering <- DistributedCache files [0] clustering, [1] clusteringCount
 *                                (M/rm/RM2Job.java:260-263; M/common/AbstractByClusterMapper.java:47-68;
 *                                 M/rm/AbstractRM2Reducer.java:93-105)
 *       fy_rm2_run            <- runUserSum + runItemColl + runItemRecommendation
 *                                (M/rm/RM2Job.java:110-151,164-205,214-270) i.e. every
 *                                AbstractRM2Reducer.reduce call (M/rm/AbstractRM2Reducer.java:129-233)
 *       fy_rm2_stats          <- rm2/userSum and rm2/itemColl outputs (M/rm/RM2Job.java:138-142,190-196)
 *       fy_rm2_results        <- every writePreference(userId, itemId, score, cluster) call
 *                                (M/rm/AbstractRM2Reducer.java:358-369,405-407; sinks
 *                                 M/rm/RM2HDFSReducer.java:44-50, M/rm/RM2CassandraReducer.java:49-63)
 *   fine seam: one AbstractRM2Reducer.reduce(key, values) call
 *       fy_rm2_score_group    <- M/rm/AbstractRM2Reducer.java:129-233 for one "c" or "c-split-nSplits" group
 *
 * Conventions
 *   - extern "C", plain pointers and sizes; callable from JNI (GetDirectBufferAddress), Panama FFM
 *     (MemorySegment) or ctypes.  No C++/torch types cross the boundary.
 *   - Every function returns FY_OK (0) or a negative fy_status; fy_rm2_last_error() gives the text.
 *     Nothing throws or aborts across the boundary (the reference's only convention is
 *     "throw RuntimeException => task fails", M/rm/RM2Job.java:265-268).
 *   - The caller owns every buffer it passes; inputs are copied during the call, never retained.
 *   - A context is single-caller; distinct contexts are independent.  One context drives one GPU, or
 *     fy_rm2_params.n_gpus of them from one process.
 *   - Capacity: one cluster needs 24 * I_c * ld bytes of device memory for its H planes (I_c = items rated in the
 *     cluster, ld = I_c rounded up to 512): 17 GB at ML-20M shape, ~86 GB at I_c = 60 000; beyond what the device
 *     holds fy_rm2_run fails with FY_E_NOMEM.
 *   - There is NO CPU fallback: without a usable CUDA device fy_rm2_create fails with FY_E_CUDA.
 *
 * Ordering contract: the reference leaves the order among equal scores to
 * java.util.PriorityQueue + hash iteration order.  This engine emits, per user, descending score
 * with ties broken by ascending item id -- the total order of the reference's Cassandra sink
 * (T/util/CassandraUtils.java:144-147).  Users are emitted grouped by (cluster, user id).
 */
#ifndef FILMYOU_RM2_H
#define FILMYOU_RM2_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FY_RM2_ABI_VERSION 3

typedef enum fy_status {
    FY_OK = 0,
    FY_E_ARG = -1,                  /* null pointer, negative size, id out of range                  */
    FY_E_USER_WITHOUT_RATING = -2,  /* clustering lists a user with no positive rating; the reducer  */
                                    /* would mis-parse the group (AbstractRM2Reducer.java:153-160)   */
    FY_E_DUPLICATE_RATING = -3,     /* same (user,item) twice: HashMap.put order is shuffle-defined  */
    FY_E_CLUSTER_SIZE = -4,         /* clusteringCount[c] != #users mapped to c                      */
    FY_E_UNKNOWN_USER = -5,         /* positive rating of a user absent from `clustering`            */
    FY_E_NOMEM = -6,                /* host or device allocation failed                              */
    FY_E_CUDA = -7,                 /* no device / CUDA runtime error (see fy_rm2_last_error)        */
    FY_E_STATE = -8,                /* call order violated (e.g. run before set_ratings)             */
    FY_E_UNSUPPORTED = -9,          /* parameter combination not implemented                         */
    FY_E_ITEM_WITHOUT_RATING = -10  /* filmyou_nmf.h: an item no user rated (WComputationMapper.java:95-98) */
} fy_status;

/* Same names / meaning as the reference's Configuration keys
 * (M/rmrecommender/RMRecommenderDriver.java:89-120; read at M/rm/AbstractRM2Reducer.java:108-110,197). */
typedef struct fy_rm2_params {
    double  lambda;            /* "lambda"                   default 0.1  (:114)                   */
    int32_t number_of_items;   /* "numberOfItems"            the configured global item count      */
    int32_t top_n;             /* "numberOfRecommendations"  default 1000 (:95)                    */
    int32_t filter_users;      /* "filterUsers"              default 0    (:119)                   */
    int32_t device;            /* CUDA device ordinal                                              */
    int32_t shard_rank;        /* this context scores shard `shard_rank` of `shard_count`          */
    int32_t shard_count;       /*   (users partitioned by estimated work; 0/1 = everything)        */
    int32_t tie_break;         /* 0 = canonical (score desc, item id asc); nothing else defined    */
    int32_t score_mode;        /* 0 = auto: stream the 4-byte hi-word plane of H, then re-score the */
                               /*     provably sufficient candidate set exactly in fp64;            */
                               /* 1 = exact: stream the fp64 plane for every term                   */
    int32_t n_gpus;            /* 0/1 = this context drives `device` only.  n > 1 = ONE context drives */
                               /*     devices device .. device+n-1 from one host process: the users are */
                               /*     sharded over them exactly as shard_rank/shard_count would (which  */
                               /*     must then be 0/1), inputs are uploaded to every device, results   */
                               /*     come back concatenated in shard order.  The GPU analogue of       */
                               /*     numReduceTasks = numberOfClusters (M/rm/RM2Job.java:251) and of    */
                               /*     the cluster-split fan-out (M/common/AbstractByClusterAndCountMapper.java:86-102) */
    int32_t reserved;          /* must be 0                                                         */
} fy_rm2_params;

typedef struct fy_rm2_ctx fy_rm2_ctx;

int fy_rm2_abi_version(void);
void fy_rm2_default_params(fy_rm2_params* p);

int fy_rm2_create(fy_rm2_ctx** out, const fy_rm2_params* p);
void fy_rm2_destroy(fy_rm2_ctx* ctx);
const char* fy_rm2_last_error(const fy_rm2_ctx* ctx);

/* Optional: run on a caller-owned CUDA stream (a cudaStream_t passed as void*), e.g. torch's
 * current stream.  Default is a stream owned by the context. */
int fy_rm2_set_stream(fy_rm2_ctx* ctx, void* cuda_stream);

/* COO ratings in any order with the reference's original ids; score <= 0 is ignored exactly as the
 * mappers do (M/rm/ScoreByClusterHDFSMapper.java:39-40).  Copies host -> device. */
int fy_rm2_set_ratings(fy_rm2_ctx* ctx, const int32_t* user, const int32_t* item, const float* score,
                       int64_t nnz);

/* `clustering` (user -> cluster, 0-based) and `clusteringCount` (cluster -> size).  n_clusters may
 * exceed the number of non-empty clusters (T/testdata/RMTestData.java:27). */
int fy_rm2_set_clustering(fy_rm2_ctx* ctx, const int32_t* user, const int32_t* cluster, int64_t n_users,
                          const int32_t* cluster_size, int32_t n_clusters);

/* The timed part: statistics, per-cluster matrices, scoring and top-N for this context's shard. */
int fy_rm2_run(fy_rm2_ctx* ctx);

/* rm2/userSum (in the order users were given to fy_rm2_set_clustering), rm2/itemColl
 * (indexed by item id, 0..max_item) and the truncated global total (M/rm/RM2Job.java:95). */
int32_t fy_rm2_max_item(const fy_rm2_ctx* ctx);
int64_t fy_rm2_user_count(const fy_rm2_ctx* ctx);      /* users given to fy_rm2_set_clustering (-1 before) */
int fy_rm2_stats(fy_rm2_ctx* ctx, double* user_sum, double* item_prob, double* total);

/* Results of the last run: packed triples grouped by (cluster, user id), descending score inside a
 * user.  Any output pointer may be NULL.  score32 is the (float) cast of the sink
 * (M/rm/RM2HDFSReducer.java:48). */
int64_t fy_rm2_result_count(const fy_rm2_ctx* ctx);
int64_t fy_rm2_users_scored(const fy_rm2_ctx* ctx);
int fy_rm2_results(fy_rm2_ctx* ctx, int32_t* user, int32_t* item, double* score64, float* score32,
                   int32_t* cluster);

/* The same results as a 12-byte-per-triple stream: fy_rm2_results(ctx, NULL, item, score64, NULL, NULL) plus one
 * (user, cluster, count) record per scored row, rows in the order of the triples.  user/cluster/score32 of a triple
 * follow from its row and (float) score64, so a host reads 12 B instead of 24 B per triple.  Any pointer may be NULL. */
int64_t fy_rm2_result_row_count(const fy_rm2_ctx* ctx);
int fy_rm2_result_rows(fy_rm2_ctx* ctx, int32_t* user, int32_t* cluster, int32_t* count);

/* Device-resident view of the same packed arrays (valid until the next run / destroy), so that a
 * multi-GPU host can hand them to NCCL without a host round trip.  Any output pointer may be NULL. */
int fy_rm2_results_device(fy_rm2_ctx* ctx, const int32_t** user, const int32_t** item, const double** score64,
                          const float** score32, const int32_t** cluster);

/* ---- one process per GPU (torchrun / MPI style): the exchange step of the job, inside the library -----------------
 * Rank 0 obtains an id with fy_rm2_nccl_unique_id and hands the bytes to the other ranks by any means; every rank then
 * calls fy_rm2_comm_init on its context (created with shard_rank = rank, shard_count = world).  From then on fy_rm2_run
 * ends with ONE grouped NCCL exchange of the dense top-N blocks over NVLink, issued on the run's stream, and
 * fy_rm2_results / result_count / result_rows on EVERY rank describe the whole job (all users, in (cluster, user id)
 * order); fy_rm2_users_scored stays this rank's share.  This replaces the shuffle + per-reducer part files of RM2-3
 * (M/rm/RM2Job.java:214-270).  libnccl.so.2 is resolved at run time (dlopen), so a process that never calls these
 * two functions does not need NCCL.  fy_rm2_comm_destroy detaches (also done by fy_rm2_destroy). */
#define FY_NCCL_UNIQUE_ID_BYTES 128
int fy_rm2_nccl_unique_id(void* id_out /* FY_NCCL_UNIQUE_ID_BYTES */);
int fy_rm2_comm_init(fy_rm2_ctx* ctx, const void* id /* FY_NCCL_UNIQUE_ID_BYTES */, int32_t world, int32_t rank);
int fy_rm2_comm_destroy(fy_rm2_ctx* ctx);
/* user-rank boundaries of the shards of the last run: bounds[shard_count + 1] */
int fy_rm2_shard_bounds(const fy_rm2_ctx* ctx, int32_t* bounds);

/* ---- f3 / north_star part 1: scoring over explicit neighbour lists ------------------------------------------------
 * buildRecommendations(..., int[] neighbours, ...) (M/rm/AbstractRM2Reducer.java:321-323,342-346) for caller-supplied
 * lists, e.g. the output of fy_knn_neighbours: user[q] is scored exactly as the reducer would score it in a group made
 * of user[q] and neighbour[q*k .. q*k+k) (-1 = empty slot): K = |N(u)| + 1, candidate items = what u or a neighbour
 * rated, neighbour sum over N(u) in ascending user id, p(i|C) / user sums = the global statistics of the ratings given
 * to fy_rm2_set_ratings (no clustering needed).  Results: fy_rm2_results / result_count / result_rows (users in the
 * order given; `cluster` = position q of the user in the call).  The reference only ever passes "the cluster minus u"
 * (:215-216): with that list this call reproduces fy_rm2_run; beyond it no reference counterpart exists.
 * Cost grows with sum_q sum_{v in {u} + N(u)} n_v expanded ratings; FY_E_UNSUPPORTED asks for a smaller user list. */
int fy_rm2_run_neighbours(fy_rm2_ctx* ctx, const int32_t* user, const int32_t* neighbour, int32_t k, int64_t n_listed);

/* Fine seam: one reduce() group, exactly the records the reducer receives
 * (M/rm/AbstractRM2Reducer.java:149-174): n_group_users (user, userSum) records, then the group's
 * rating records; item_prob is rm2/itemColl indexed by item id (size max_item+1).
 * Only users with user % n_splits == split are scored (:203-205).  Results via fy_rm2_results. */
int fy_rm2_score_group(fy_rm2_ctx* ctx, int32_t cluster_id, int32_t split, int32_t n_splits,
                       const int32_t* group_user, const double* group_user_sum, int32_t n_group_users,
                       const int32_t* r_user, const int32_t* r_item, const float* r_score, int64_t nnz,
                       const double* item_prob, int32_t max_item);

/* Work and timing of the last run, for the bench (all measured with CUDA events on the stream the
 * kernels were launched on). */
typedef struct fy_rm2_profile {
    double ms_total;        /* whole fy_rm2_run                                                  */
    double ms_index;        /* sort / CSR / CSC / statistics                                     */
    double ms_gram;         /* per-cluster H = S + b a^T build                                   */
    double ms_score;        /* the streaming score kernel (dominant)                             */
    double ms_topn;         /* radix select + sort                                               */
    double log_terms;       /* sum over scored users of n_u * I_c  (matrix elements streamed)    */
    double score_bytes;     /* algorithmic bytes of the score kernel = 8 * log_terms             */
    double gram_bytes;      /* bytes of H written                                                */
    int64_t users_scored;
    int64_t kernel_launches; /* launches of this library's own kernels inside fy_rm2_run         */
    int32_t clusters_touched;
    int32_t score_launches;
    double ms_refine;       /* margin gather + exact fp64 re-score of the candidates (auto mode) */
    double bytes_per_term;  /* 4 (hi-word plane) or 8 (fp64 plane): what score_bytes counts      */
    int32_t exact_rerun;    /* 1 if a candidate list overflowed and the run was redone in exact mode */
    int32_t score_kernel;   /* dominant kernel of the last run: 0 = k_score (fp64), 1 = k_score_hi, 2 = k_score_f32 */
    double ms_gather;       /* the NCCL exchange of the top-N blocks (0 without a communicator)      */
} fy_rm2_profile;
int fy_rm2_get_profile(const fy_rm2_ctx* ctx, fy_rm2_profile* out);

/* Roofline probe for the bench: the memory side of the dominant score kernel with the arithmetic removed (same grid,
 * same 16-byte loads, `rows_per_user` pseudo-random rows of an [n_rows x ld] 4-byte plane per (user, 512-column tile)
 * CTA).  Returns the GB/s the L2 -> SM path of THIS box delivers for that access pattern. */
int fy_rm2_probe_plane_read(fy_rm2_ctx* ctx, int32_t n_rows, int32_t n_users, int32_t rows_per_user, int32_t reps,
                            double* gb_per_s, double* ms_per_launch);

/* ---- Config 3: item-item co-occurrence counts on the binarised ratings -------
 * Replaces the Mahout RowSimilarityJob(CooccurrenceCountSimilarity) call at
 * M/baselinerecommender/BaselineRecommenderJob.java:241-253.  C[i][j] = #users with both items,
 * int32, dense [n_items x n_items] over item ids 0..n_items-1 (bit-exact integer counts; int8
 * tcgen05 GEMM with int32 TMEM accumulators).  PARITY UNPINNED: Mahout 0.8 is not vendored. */
int fy_cooc_counts(fy_rm2_ctx* ctx, int32_t n_user_ids, int32_t n_items, int32_t* counts_out /* host, may be NULL */,
                   double* ms_gemm_out);
/* per-row top-k of the last fy_cooc_counts, self excluded (maxSimilaritiesPerRow=100,
 * excludeSelfSimilarity=true, BaselineRecommenderJob.java:247-250); ties by ascending item id. */
int fy_cooc_topk(fy_rm2_ctx* ctx, int32_t k, int32_t* item_out /* [n_items*k] */, int32_t* count_out /* [n_items*k] */,
                 int32_t* n_out /* [n_items] */);

/* ---- a9 / f3: kNN neighbourhood provider (north_star part 1; NO reference symbol exists) ----------
 * user-user co-occurrence counts on the binarised ratings (int8 tcgen05 GEMM in row blocks) and, per
 * user id, the k users sharing most items (self excluded, ties by ascending user id, users with no
 * common item omitted).  Checked against an integer CPU restatement only; not used by fy_rm2_run, whose
 * neighbourhood is the reference's: the user's cluster (M/rm/AbstractRM2Reducer.java:215-216). */
int fy_knn_neighbours(fy_rm2_ctx* ctx, int32_t n_user_ids, int32_t n_items, int32_t k,
                      int32_t* neighbour_out /* [n_user_ids*k], -1 padded */, int32_t* count_out /* may be NULL */,
                      int32_t* n_out /* [n_user_ids], may be NULL */, double* ms_gemm_out /* may be NULL */);

#ifdef __cplusplus
}
#endif
#endif
