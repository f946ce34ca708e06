/*
 * filmyou_nmf.h -- C ABI of the B200-native NMF / PPC clustering step (libfilmyou_rm2.so).
 *
 * the job chain that PRODUCES the `clustering` / `clusteringCount` files the RM2
 * hot path consumes (filmyou_rm2.h).  M/ = /root/reference/src/main/java/es/udc/fi/dc/irlab/.
 *
 *   fy_nmf_set_ratings   <- the ratings SequenceFile read by VectorByItemHDFSMapper / ItemScoreByUserHDFSMapper
 *                           (M/nmf/hcomputation/VectorByItemHDFSMapper.java:36-58,
 *                            M/nmf/wcomputation/ItemScoreByUserHDFSMapper.java:33-54)
 *   fy_nmf_set_factors   <- the H / W SequenceFile<IntWritable, VectorWritable> named by the "H" / "W"
 *                           options (M/nmf/AbstractNMFDriver.java:102-106)
 *   fy_nmf_init_random   <- createInitialMatrices (M/nmf/AbstractNMFDriver.java:79-88,
 *                           M/util/DataInitialization.java:68-95: uniform + eps, rows L1-normalised)
 *   fy_nmf_run           <- numberOfIterations x (ComputeHJob | PPCComputeHJob, ComputeWJob)
 *                           (M/nmf/AbstractNMFDriver.java:117-138; 4 + 5 MapReduce jobs per iteration:
 *                            M/nmf/hcomputation/ComputeHJob.java:74-101, M/nmf/wcomputation/ComputeWJob.java:72-98)
 *   fy_nmf_get_factors   <- the H / W files after the last iteration
 *   fy_nmf_cluster_assignment <- ClusterAssignmentJob + CountClustersJob
 *                           (M/nmf/clustering/FindClusterMapper.java:34-42, M/nmf/clustering/CountReducer.java:35-45)
 *
 * Same conventions as filmyou_rm2.h: extern "C", plain pointers, fy_status return codes, inputs copied
 * during the call, no CPU fallback.  All arithmetic is fp64 with separately rounded multiply and add
 * (Java has no FMA contraction), in a fixed summation order -- see fy_nmf_params.combine_len.
 */
#ifndef FILMYOU_NMF_H
#define FILMYOU_NMF_H

#include <stdint.h>
#include "filmyou_rm2.h"   /* fy_status */

#ifdef __cplusplus
extern "C" {
#endif

/* Names follow the reference's Configuration keys (M/rmrecommender/RMRecommenderDriver.java:89-120). */
typedef struct fy_nmf_params {
    int32_t mode;                    /* 0 = NMFDriver (M/nmf/NMFDriver.java), 1 = PPCDriver (M/nmf/ppc/PPCDriver.java) */
    int32_t number_of_users;         /* "numberOfUsers": rows of H, user ids id_base .. id_base+U-1          */
    int32_t number_of_items;         /* "numberOfItems": rows of W                                            */
    int32_t number_of_clusters;      /* "numberOfClusters": columns of H and W (<= 512 here)                 */
    int32_t number_of_iterations;    /* "numberOfIterations", default 10 (:94)                               */
    int32_t normalization_frequency; /* "normalizationFrequency", default 12 (:115)                          */
    int32_t apply_normalization;     /* 0 = what the reference DOES: PPCHComputationReducer.java:88-90 drops  */
                                     /* the vector normalize(1) returns, so rows are never renormalised       */
                                     /* (golden-checked); 1 = apply the intended L1 renormalisation            */
    int32_t id_base;                 /* first user / item id, 1 in the reference's files                      */
    int32_t combine_len;             /* ratings per combiner group of the join sums (VectorSumReducer runs as */
                                     /* combiner and reducer); group sums are added in ascending order.       */
                                     /* default 1024; 0 = one sequential sum per row                          */
    int32_t split_rows;              /* rows per combiner group of the cross-product sums (MatrixSumReducer), */
                                     /* default 256; 0 = one sequential sum                                    */
    int32_t device;                  /* CUDA device ordinal                                                    */
} fy_nmf_params;

typedef struct fy_nmf_ctx fy_nmf_ctx;

void fy_nmf_default_params(fy_nmf_params* p);
int fy_nmf_create(fy_nmf_ctx** out, const fy_nmf_params* p);
void fy_nmf_destroy(fy_nmf_ctx* ctx);
const char* fy_nmf_last_error(const fy_nmf_ctx* ctx);

/* COO ratings, any order, ids in [id_base, id_base + n); score <= 0 is dropped as the mappers do.
 * A user (item) left without a positive rating fails the run with FY_E_USER_WITHOUT_RATING
 * (FY_E_ITEM_WITHOUT_RATING), as HComputationReducer.java:52-55 / WComputationMapper.java:95-98 throw. */
int fy_nmf_set_ratings(fy_nmf_ctx* ctx, const int32_t* user, const int32_t* item, const float* score, int64_t nnz);

/* H [number_of_users x k], W [number_of_items x k], row-major doubles; row r = id id_base + r. */
int fy_nmf_set_factors(fy_nmf_ctx* ctx, const double* H, const double* W);
/* Random start as createMatrix builds it; the reference seeds java.util.Random from the clock, so only
 * the distribution is reproduced (counter-based generator, same `seed` -> same matrices). */
int fy_nmf_init_random(fy_nmf_ctx* ctx, uint64_t seed);

int fy_nmf_run(fy_nmf_ctx* ctx);

int fy_nmf_get_factors(fy_nmf_ctx* ctx, double* H, double* W);   /* either may be NULL */

/* cluster_out[r] = arg-max of row r of H (Mahout maxValueIndex: first strict maximum);
 * cluster_size_out[c] = users assigned to c.  Feed both to fy_rm2_set_clustering. */
int fy_nmf_cluster_assignment(fy_nmf_ctx* ctx, int32_t* cluster_out, int32_t* cluster_size_out);

typedef struct fy_nmf_profile {
    double ms_index;          /* sorts, CSR / CSC, combiner segments (once per fy_nmf_run)            */
    double ms_iterations;     /* all iterations, CUDA events on the launching stream                  */
    double ms_per_iteration;
    double join_bytes;        /* algorithmic bytes of the two join sums per iteration:                */
                              /*   nnz * (8 k + 12) * 2  (one factor row gathered + key + score)      */
    int64_t ratings;          /* positive ratings indexed                                             */
    int64_t kernel_launches;  /* this library's kernels launched by the last fy_nmf_run               */
    int32_t iterations;
    int32_t graph_replays;    /* iterations replayed from the captured CUDA graph                     */
} fy_nmf_profile;
int fy_nmf_get_profile(const fy_nmf_ctx* ctx, fy_nmf_profile* out);

#ifdef __cplusplus
}
#endif
#endif
