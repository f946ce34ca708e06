// rm2_kernels.cuh -- device code of the RM2 engine (sm_100a).
//
// What the reference does per cluster (M/rm/AbstractRM2Reducer.java:129-233,321-371):
//   P[v][i] = (1-lambda)*(r_vi/S_v) + lambda*p(i|C)                       (:384-389)
//   score(u,i) = pvpi(u) + sum_{j in rated(u)} log( sum_{v != u} P[v][i]*P[v][j] )   (:332-354)
//   top-N by descending score                                             (:358-369)
//
// How this engine computes the same numbers.  Write P[v][i] = b_i + d_vi with b_i = lambda*p(i|C)
// (the exact value of P for an unrated cell) and d_vi = P[v][i] - b_i (non-zero only where v rated
// i).  For a candidate i (unrated by u) and a rated item j of u,
//   t(u,i,j) = sum_{v != u} P[v][i]*P[v][j] = H[j][i] + b_i * c(u,j)
//   H[j][i]  = S[j][i] + b_j * alpha_i,   S[j][i] = sum_{v rated i and j} d_vi*d_vj,
//   alpha_i  = sum_{v rated i} d_vi,      c(u,j)  = (K-1)*b_j + sum_{v != u, v rated j} d_vj.
// Every term is non-negative, so nothing cancels: t carries a few ulp of error for any lambda.
// S is built from the sparse ratings (work sum_v n_v^2, not the dense 2*I^2*K of a GEMM), H is
// written once per cluster (I_c^2 doubles) and the score kernel streams row H[j][.] for every
// rated j of every user: 8 bytes per (u,i,j) log-term, the HBM-bound part.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace fy {

// -DFY_BOUNDS_CHECK: every index into a shared-memory accumulator, a candidate list, a staged row list or an output
// block is range-checked on the device and counted (fy_rm2_debug_violations()).  compute-sanitizer is closed on the
// development pool (profiles/r02_sanitizer_closed.txt), so the GPU suite is run once against a library built this way
// (tools/build_checked.py); the product build compiles the checks out.
#ifdef FY_BOUNDS_CHECK
__device__ unsigned long long g_bounds_violations = 0ull;
#define FY_CHECK(cond) do { if (!(cond)) atomicAdd(&fy::g_bounds_violations, 1ull); } while (0)
#else
#define FY_CHECK(cond) do { } while (0)
#endif

// error flags raised on the device, read back at the host sync points
enum DevFlag : int {
    DF_UNKNOWN_USER = 0,
    DF_DUPLICATE = 1,
    DF_USER_WITHOUT_RATING = 2,
    DF_BAD_ITEM = 3,
    DF_INEXACT_SCORES = 4,   // informational: some score is not a multiple of 2^-16 below 2^15
    DF_COUNT = 8
};

constexpr int SCORE_TILE = 256;     // candidates per score CTA (2 per thread)
constexpr int SCORE_THREADS = 128;
constexpr int SCORE_CHUNK = 512;    // rated items staged in shared memory at a time
constexpr int SCOREH_TILE = 512;    // candidates per CTA of the hi-word score kernel (4 per thread)
constexpr int REFINE_THREADS = 256;
#ifndef FY_H_SLICE_COLS
#define FY_H_SLICE_COLS 512
#endif
constexpr int H_SLICE = FY_H_SLICE_COLS;   // columns of one H row built by one warp (4 KB of doubles)
constexpr int H_THREADS = 256;
constexpr int H_WARPS = H_THREADS / 32;   // 8 independent (row, slice) tasks per CTA
constexpr int TOPN_THREADS = 512;
constexpr int TOPN_MAX_SELECT = 4096;

// ---------------------------------------------------------------------------------------------
// Indexing
// ---------------------------------------------------------------------------------------------

// rating -> sort key (rank << item_bits | item); ratings with score <= 0 get the all-ones key and
// sort to the end (the mappers drop them: M/rm/ScoreByClusterHDFSMapper.java:39-40).
__global__ void k_make_keys(const int32_t* __restrict__ r_user, const int32_t* __restrict__ r_item,
                            const float* __restrict__ r_score, int64_t nnz,
                            const int32_t* __restrict__ uid_sorted, const int32_t* __restrict__ uid_rank,
                            int32_t n_users, const int32_t* __restrict__ uid_table, int32_t uid_table_n, int32_t uid_table_min,
                            int item_bits, int32_t max_item_allowed,
                            uint64_t* __restrict__ keys, unsigned long long* __restrict__ n_valid,
                            int* __restrict__ flags,
                            // optional (sharded index, exact scores): the global statistics in the same pass
                            const int32_t* __restrict__ rank_cluster, int32_t table_items, double* __restrict__ usum,
                            int32_t* __restrict__ n_u, double* __restrict__ isum, int32_t* __restrict__ present) {
    int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int local_valid = 0;
    if (e < nnz) {
        uint64_t key = ~0ull;
        if (r_score[e] > 0.0f) {
            const int32_t uid = r_user[e];
            int rank = -1;
            if (uid_table) {                                  // dense ids: one load
                const int64_t o = (int64_t)uid - (int64_t)uid_table_min;
                if (o >= 0 && o < uid_table_n) rank = uid_table[o];
            } else {
                int lo = 0, hi = n_users - 1;
                while (lo <= hi) {
                    const int mid = (lo + hi) >> 1;
                    const int32_t v = uid_sorted[mid];
                    if (v == uid) { rank = uid_rank[mid]; break; }
                    if (v < uid) lo = mid + 1; else hi = mid - 1;
                }
            }
            const int32_t it = r_item[e];
            if (rank < 0) {
                atomicOr(&flags[DF_UNKNOWN_USER], 1);
            } else if (it < 0 || it > max_item_allowed) {
                atomicOr(&flags[DF_BAD_ITEM], 1);
            } else {
                key = ((uint64_t)(uint32_t)rank << item_bits) | (uint64_t)(uint32_t)it;
                local_valid = 1;
                if (usum) {
                    const double sc = (double)r_score[e];
                    atomicAdd(&usum[rank], sc);      // exact for dyadic scores, hence order-independent (DF_INEXACT_SCORES clear)
                    atomicAdd(&isum[it], sc);
                    atomicAdd(&n_u[rank], 1);
                    present[(size_t)rank_cluster[rank] * table_items + it] = 0;      // table pre-set to -1
                }
            }
        }
        keys[e] = key;
    }
    const int block_valid = __syncthreads_count(local_valid);
    if (threadIdx.x == 0 && block_valid) atomicAdd(n_valid, (unsigned long long)block_valid);
}

// set_ratings-time scan: largest item id among positive ratings, negative ids flagged
__global__ void k_scan_ratings(const int32_t* __restrict__ r_item, const float* __restrict__ r_score, int64_t nnz,
                               int* __restrict__ max_item, int* __restrict__ flags) {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int local_max = -1;
    if (e < nnz && r_score[e] > 0.0f) {
        local_max = r_item[e];
        if (local_max < 0) { atomicOr(&flags[DF_BAD_ITEM], 1); local_max = -1; }
        // sums of such scores are exact in double in ANY order (up to 2^22 addends), which lets the
        // item sums be reduced in parallel and still equal the reference's sequential double sum
        const float sc = r_score[e] * 65536.0f;
        if (!(r_score[e] < 32768.0f) || sc != truncf(sc)) atomicOr(&flags[DF_INEXACT_SCORES], 1);
    }
    for (int o = 16; o > 0; o >>= 1) local_max = max(local_max, __shfl_xor_sync(0xffffffffu, local_max, o));
    if ((threadIdx.x & 31) == 0 && local_max >= 0) atomicMax(max_item, local_max);
}

// ---- sharded index (shard_count > 1, exact-score inputs): global statistics without a sort, then only
// the ratings of the clusters this rank touches are compacted, sorted and indexed ----
__global__ void k_global_stats(const uint64_t* __restrict__ keys, const float* __restrict__ score, int64_t nnz,
                               int item_bits, const int32_t* __restrict__ rank_cluster, int32_t table_items,
                               double* __restrict__ usum, int32_t* __restrict__ n_u, double* __restrict__ isum,
                               int32_t* __restrict__ present) {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= nnz) return;
    const uint64_t k = keys[e];
    if (k == ~0ull) return;
    const int32_t rank = (int32_t)(k >> item_bits);
    const int32_t item = (int32_t)(k & ((1ull << item_bits) - 1));
    const double s = (double)score[e];
    atomicAdd(&usum[rank], s);          // exact for dyadic scores, hence order-independent (DF_INEXACT_SCORES clear)
    atomicAdd(&isum[item], s);
    atomicAdd(&n_u[rank], 1);
    present[(size_t)rank_cluster[rank] * table_items + item] = 0;      // table pre-set to -1
}

__global__ void k_total_from_usum(const double* __restrict__ usum, const int32_t* __restrict__ n_u, int32_t n_users,
                                  unsigned long long* __restrict__ counter, int* __restrict__ flags) {
    const int32_t u = blockIdx.x * blockDim.x + threadIdx.x;
    if (u >= n_users) return;
    if (n_u[u] == 0) { atomicOr(&flags[DF_USER_WITHOUT_RATING], 1); return; }
    atomicAdd(counter, (unsigned long long)((long long)usum[u]) * 100ull);   // DoubleSumAndCountReducer.java:41
}

__global__ void k_user_work_n(const int32_t* __restrict__ n_u, const int32_t* __restrict__ rank_cluster,
                              const int32_t* __restrict__ icount, int32_t n_users, double* __restrict__ work) {
    const int32_t u = blockIdx.x * blockDim.x + threadIdx.x;
    if (u >= n_users) return;
    work[u] = (double)n_u[u] * (double)icount[rank_cluster[u]];
}

__global__ void k_item_prob_isum(const double* __restrict__ isum, int32_t table_items,
                                 const unsigned long long* __restrict__ counter, double lambda,
                                 double* __restrict__ iprob, double* __restrict__ bvec,
                                 double* __restrict__ total_out, unsigned long long* __restrict__ bmin_bits) {
    const int32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    const double total = __ddiv_rn((double)(long long)(*counter), 100.0);
    if (i == 0) *total_out = total;
    if (i >= table_items) return;
    const double s = isum[i];
    const double p = (s > 0.0) ? __ddiv_rn(s, total) : 0.0;
    const double b = __dmul_rn(lambda, p);
    iprob[i] = p; bvec[i] = b;
    if (s > 0.0) atomicMin(bmin_bits, (unsigned long long)__double_as_longlong(b > 0.0 ? b : 0.0));
}

// Work partition over `world` ranks.  A rank pays the score work of its users (n_u * I_c per user, W[c] per whole cluster)
// PLUS, once per cluster it touches, that cluster's H build (~ beta * W[c]: the build of a cluster costs about beta times
// its score + top-N + re-score work at both benchmark shapes) -- the clusters that straddle a boundary are built twice,
// so equal score work alone leaves the ranks with one more touched cluster ~10 % late (measured at 8 GPUs).  The minimal
// maximum cost T is found by a 33-section search (8 rounds of 32 probes) over a greedy fill (contiguous ranges, monotone
// cost), G[r] = cumulative score work at the end of rank r - 1.  Pure double arithmetic in a fixed order with separately
// rounded operations: device, host and the Python mirror agree.
// greedy fill with budget T per rank: G[r + 1] = cumulative score work after rank r; true when everything is covered
__host__ __device__ inline bool partition_fill(const double* W, int n_clusters, double beta, int world, double T, double* G) {
    int c = 0;
    while (c < n_clusters && !(W[c] > 0.0)) c++;
    double rem = (c < n_clusters) ? W[c] : 0.0, pos = 0.0;
    if (G) G[0] = 0.0;
    for (int r = 0; r < world; r++) {
        double budget = T;
        while (c < n_clusters) {
#ifdef __CUDA_ARCH__
            const double fee = __dmul_rn(beta, W[c]);
#else
            const double fee = beta * W[c];
#endif
            if (!(budget > fee)) break;
            budget -= fee;
            const double take = rem < budget ? rem : budget;
            rem -= take; budget -= take; pos += take;
            if (rem > 0.0) break;
            c++;
            while (c < n_clusters && !(W[c] > 0.0)) c++;
            rem = (c < n_clusters) ? W[c] : 0.0;
        }
        if (G) G[r + 1] = pos;
    }
    return c >= n_clusters;
}
// the l-th of 32 probe points strictly inside (lo, hi); separate roundings, so that device, host and Python agree
__host__ __device__ inline double partition_probe(double lo, double hi, int l) {
#ifdef __CUDA_ARCH__
    return __dadd_rn(lo, __ddiv_rn(__dmul_rn(__dsub_rn(hi, lo), (double)(l + 1)), 33.0));
#else
    const double w = hi - lo;
    const double x = w * (double)(l + 1);
    return lo + x / 33.0;
#endif
}
constexpr int PARTITION_ROUNDS = 8;                  // 33^-8 = 7e-13 of the initial bracket

// host form of the search (the device form below evaluates the 32 probes of a round on the 32 lanes of a warp)
inline void partition_targets(const double* W, int n_clusters, double beta, int world, double* G /* [world + 1] */) {
    double total = 0.0;
    for (int c = 0; c < n_clusters; c++) total += W[c];
    double lo = 0.0, hi = total * (1.0 + beta) + 1.0;
    for (int round = 0; round < PARTITION_ROUNDS; round++) {
        int first = 32;
        for (int l = 0; l < 32; l++)
            if (partition_fill(W, n_clusters, beta, world, partition_probe(lo, hi, l), nullptr)) { first = l; break; }
        const double nlo = first > 0 ? partition_probe(lo, hi, first - 1) : lo;
        const double nhi = first < 32 ? partition_probe(lo, hi, first) : hi;
        lo = nlo; hi = nhi;
    }
    partition_fill(W, n_clusters, beta, world, hi, G);
    G[world] = total;
}

constexpr int MAX_PARTITION_CLUSTERS = 4096;      // beyond this (or world > 64) the plain equal-work rule is used

// Shard boundaries on the device (no host round trip): shard r = user ranks [bounds[r], bounds[r+1]); then the rank
// range [rl, rh) of the clusters this rank touches.  out = bounds[world + 1], rl, rh.  scan = inclusive prefix of the
// per-user work (integers below 2^53, so every rank computes the same numbers).
__global__ void k_shard_bounds(const double* __restrict__ scan, int32_t n_users, int32_t world, int32_t me,
                               const int32_t* __restrict__ rank_cluster, const int32_t* __restrict__ cstart, int32_t n_clusters,
                               double beta, int32_t* __restrict__ out) {
    __shared__ double s_W[MAX_PARTITION_CLUSTERS];
    __shared__ double s_G[66];
    const bool aware = n_clusters <= MAX_PARTITION_CLUSTERS && world <= 64 && beta > 0.0;
    if (aware) {
        for (int c = threadIdx.x; c < n_clusters; c += blockDim.x) {
            const int32_t a = cstart[c], b = cstart[c + 1];
            s_W[c] = (b > a) ? scan[b - 1] - (a > 0 ? scan[a - 1] : 0.0) : 0.0;
        }
        __syncthreads();
        if (threadIdx.x < 32) {                                   // one warp: the 32 probes of a round on its 32 lanes
            const int l = threadIdx.x;
            double total = 0.0;
            for (int c = 0; c < n_clusters; c++) total += s_W[c];
            double lo = 0.0, hi = __dadd_rn(__dmul_rn(total, __dadd_rn(1.0, beta)), 1.0);
            for (int round = 0; round < PARTITION_ROUNDS; round++) {
                const bool ok = partition_fill(s_W, n_clusters, beta, world, partition_probe(lo, hi, l), nullptr);
                const unsigned bal = __ballot_sync(0xffffffffu, ok);
                const int first = bal ? (__ffs(bal) - 1) : 32;
                const double nlo = first > 0 ? partition_probe(lo, hi, first - 1) : lo;
                const double nhi = first < 32 ? partition_probe(lo, hi, first) : hi;
                lo = nlo; hi = nhi;
            }
            if (l == 0) { partition_fill(s_W, n_clusters, beta, world, hi, s_G); s_G[world] = total; }
        }
        __syncthreads();
    }
    const int r = threadIdx.x;
    if (r <= world) {
        int32_t b;
        if (r == 0) b = 0;
        else if (r >= world) b = n_users;
        else {
            const double target = aware ? s_G[r] : scan[n_users - 1] * (double)r / (double)world;
            int32_t lo = 0, hi = n_users;
            while (lo < hi) { const int32_t mid = (lo + hi) >> 1; if (scan[mid] < target) lo = mid + 1; else hi = mid; }
            b = lo;
            if (aware && b < n_users) {
                // a target that is the end of a cluster (up to rounding) must fall ON the cluster boundary: one stray user
                // on the wrong side would make a rank build a whole extra H
                const double tol = 1e-9 * scan[n_users - 1];
                const int32_t c = rank_cluster[b];
                const int32_t e0 = cstart[c], e1 = cstart[c + 1];
                const double p0 = e0 > 0 ? scan[e0 - 1] : 0.0, p1 = scan[e1 - 1];
                if (fabs(p0 - target) <= tol) b = e0;
                else if (fabs(p1 - target) <= tol) b = e1;
            }
        }
        out[r] = b;
    }
    __syncthreads();
    if (r == 0) {
        for (int q = 1; q <= world; q++) if (out[q] < out[q - 1]) out[q] = out[q - 1];       // monotone
        const int32_t ub = out[me], ue = out[me + 1];
        int32_t rl = 0, rh = 0;
        if (ue > ub) { rl = cstart[rank_cluster[ub]]; rh = cstart[rank_cluster[ue - 1] + 1]; }
        out[world + 1] = rl; out[world + 2] = rh;
    }
}

// keep the ratings whose user rank lies in [rl, rh) (the clusters this rank touches); warp-aggregated append
__global__ void k_compact_local(const uint64_t* __restrict__ keys, const float* __restrict__ score, int64_t nnz,
                                int item_bits, const int32_t* __restrict__ rlrh, uint64_t* __restrict__ keys_out,
                                float* __restrict__ score_out, unsigned long long* __restrict__ counter) {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int32_t rl = rlrh[0], rh = rlrh[1];
    bool keep = false;
    uint64_t k = 0;
    if (e < nnz) {
        k = keys[e];
        if (k != ~0ull) { const int32_t rank = (int32_t)(k >> item_bits); keep = (rank >= rl && rank < rh); }
    }
    const unsigned bal = __ballot_sync(0xffffffffu, keep);
    if (bal == 0u) return;
    const int lane = threadIdx.x & 31;
    unsigned long long base = 0;
    if (lane == (__ffs(bal) - 1)) base = atomicAdd(counter, (unsigned long long)__popc(bal));
    base = __shfl_sync(0xffffffffu, base, __ffs(bal) - 1);
    if (keep) {
        const unsigned long long pos = base + (unsigned long long)__popc(bal & ((1u << lane) - 1));
        FY_CHECK(pos < (unsigned long long)nnz);
        keys_out[pos] = k;
        score_out[pos] = score[e];
    }
}

__global__ void k_init_cbound(unsigned long long* __restrict__ cbound, int32_t n_clusters) {
    const int32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c < n_clusters) { cbound[3 * c] = 0ull; cbound[3 * c + 1] = 0ull; cbound[3 * c + 2] = ~0ull; }
}

// users with at least one emitted item and the log-terms of the scored range: out[0] += users, dout[0] += terms
__global__ void k_run_totals(const int32_t* __restrict__ out_count, const double* __restrict__ work, int32_t n_rows,
                             unsigned long long* __restrict__ users, double* __restrict__ terms) {
    const int32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    int scored = 0;
    double t = 0.0;
    if (r < n_rows) { scored = out_count[r] > 0; t = work[r]; }
    for (int o = 16; o > 0; o >>= 1) { scored += __shfl_xor_sync(0xffffffffu, scored, o); t += __shfl_xor_sync(0xffffffffu, t, o); }
    if ((threadIdx.x & 31) == 0) { if (scored) atomicAdd(users, (unsigned long long)scored); atomicAdd(terms, t); }   // integers < 2^53: exact in any order
}

// sorted keys -> CSR row pointers over user ranks (+ duplicate detection)
__global__ void k_rows(const uint64_t* __restrict__ keys, int32_t m, int item_bits, int32_t n_users,
                       int32_t* __restrict__ rowptr, int* __restrict__ flags) {
    const int32_t e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= m) return;
    const uint64_t k = keys[e];
    const int32_t r = (int32_t)(k >> item_bits);
    int32_t prev = -1;
    if (e > 0) {
        const uint64_t kp = keys[e - 1];
        prev = (int32_t)(kp >> item_bits);
        if (kp == k) atomicOr(&flags[DF_DUPLICATE], 1);
    }
    if (r - prev <= 64) {
        for (int32_t q = prev + 1; q <= r; q++) rowptr[q] = e;
    } else {
        // long gap (ranks of other shards): leave it to k_rows_gap, which fills it in parallel
        rowptr[r] = e;
    }
}

// fills the rank gaps k_rows skipped: every rank whose rowptr is still "unset" (-1) takes the value of
// the next set entry (rowptr[n_users] = m is set by the host memset + this kernel's boundary rule)
__global__ void k_rows_gap(const uint64_t* __restrict__ keys, int32_t m, int item_bits, int32_t n_users,
                           int32_t* __restrict__ rowptr) {
    const int32_t q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q > n_users) return;
    if (rowptr[q] >= 0) return;
    // lower_bound of rank q in the sorted keys
    int32_t a = 0, b = m;
    while (a < b) { const int32_t mid = (a + b) >> 1; if ((int32_t)(keys[mid] >> item_bits) < q) a = mid + 1; else b = mid; }
    rowptr[q] = a;
}

// RM2-1: user sums in ascending item order (M/rm/DoubleSumReducer.java:31-42) and the truncated
// counter += (long) sum * 100 (M/rm/DoubleSumAndCountReducer.java:41).
__global__ void k_user_sum(const int32_t* __restrict__ rowptr, const float* __restrict__ s_score,
                           int32_t n_users, const double* __restrict__ ext_usum,
                           double* __restrict__ usum, unsigned long long* __restrict__ counter,
                           int* __restrict__ flags) {
    const int32_t u = blockIdx.x * blockDim.x + threadIdx.x;
    if (u >= n_users) return;
    const int32_t e0 = rowptr[u], e1 = rowptr[u + 1];
    if (e1 <= e0) { atomicOr(&flags[DF_USER_WITHOUT_RATING], 1); usum[u] = 0.0; return; }
    double s = 0.0;
    for (int32_t e = e0; e < e1; e++) s = __dadd_rn(s, (double)s_score[e]);
    if (ext_usum) s = ext_usum[u];
    usum[u] = s;
    atomicAdd(counter, (unsigned long long)((long long)s) * 100ull);
}

// second sort key: item << rank_bits | rank  (CSC order: item-major, then (cluster,user) rank)
__global__ void k_make_keys2(const uint64_t* __restrict__ keys, int32_t m, int item_bits, int rank_bits,
                             uint64_t* __restrict__ keys2, int32_t* __restrict__ src) {
    const int32_t e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= m) return;
    const uint64_t k = keys[e];
    const uint64_t item = k & ((1ull << item_bits) - 1);
    const uint64_t rank = k >> item_bits;
    keys2[e] = (item << rank_bits) | rank;
    src[e] = e;
}

// CSC boundaries: per item [ifirst, ilast) and per (cluster, item) group [tstart, tend)
__global__ void k_item_groups(const uint64_t* __restrict__ keys2, int32_t m, int rank_bits,
                              const int32_t* __restrict__ rank_cluster, int32_t table_items,
                              int32_t* __restrict__ ifirst, int32_t* __restrict__ ilast,
                              int32_t* __restrict__ tstart, int32_t* __restrict__ tend) {
    const int32_t x = blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= m) return;
    const uint64_t rmask = (1ull << rank_bits) - 1;
    const uint64_t k = keys2[x];
    const int32_t item = (int32_t)(k >> rank_bits);
    const int32_t c = rank_cluster[(int32_t)(k & rmask)];
    int32_t pitem = -1, pc = -1, nitem = -1, nc = -1;
    if (x > 0) { const uint64_t kp = keys2[x - 1]; pitem = (int32_t)(kp >> rank_bits); pc = rank_cluster[(int32_t)(kp & rmask)]; }
    if (x < m - 1) { const uint64_t kn = keys2[x + 1]; nitem = (int32_t)(kn >> rank_bits); nc = rank_cluster[(int32_t)(kn & rmask)]; }
    const size_t t = (size_t)c * table_items + item;
    if (item != pitem) ifirst[item] = x;
    if (item != nitem) ilast[item] = x + 1;
    if (item != pitem || c != pc) tstart[t] = x;
    if (item != nitem || c != nc) tend[t] = x + 1;
}

// RM2-2: item sums in ascending (cluster,user) order, p(i|C) = sum / total
// (M/rm/DoubleSumAndDividerReducer.java:32-46), b_i = lambda * p(i|C).
__global__ void k_item_prob(const int32_t* __restrict__ ifirst, const int32_t* __restrict__ ilast,
                            const int32_t* __restrict__ csc_src, const float* __restrict__ s_score,
                            int32_t table_items, const unsigned long long* __restrict__ counter,
                            const double* __restrict__ ext_iprob, double lambda,
                            double* __restrict__ isum, double* __restrict__ iprob, double* __restrict__ bvec,
                            double* __restrict__ total_out, unsigned long long* __restrict__ bmin_bits) {
    const int32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    const double total = __ddiv_rn((double)(long long)(*counter), 100.0);   // RM2Job.java:95
    if (i == 0) *total_out = total;
    if (i >= table_items) return;
    const int32_t x0 = ifirst[i], x1 = ilast[i];
    double s = 0.0;
    for (int32_t x = x0; x < x1; x++) s = __dadd_rn(s, (double)s_score[csc_src[x]]);
    double p = (x1 > x0) ? __ddiv_rn(s, total) : 0.0;
    if (ext_iprob) p = ext_iprob[i];
    isum[i] = s;
    iprob[i] = p;
    const double b = __dmul_rn(lambda, p);
    bvec[i] = b;
    // smallest b over rated items (positive doubles order like their bit patterns); 0 if any b is 0
    if (x1 > x0) atomicMin(bmin_bits, (unsigned long long)__double_as_longlong(b > 0.0 ? b : 0.0));
}

// fast path of RM2-2 when every score is a small dyadic rational (DF_INEXACT_SCORES clear): the
// double sums are exact, so partial sums per (cluster, item) group combined per item give the same
// bits as the sequential sum above.
__global__ void k_group_sum(const int32_t* __restrict__ tstart, const int32_t* __restrict__ tend,
                            const int32_t* __restrict__ csc_src, const float* __restrict__ s_score,
                            size_t n_cells, double* __restrict__ tsum) {
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_cells) return;
    const int32_t x0 = tstart[t];
    double s = 0.0;
    if (x0 >= 0) {
        const int32_t x1 = tend[t];
        for (int32_t x = x0; x < x1; x++) s += (double)s_score[csc_src[x]];
    }
    tsum[t] = s;
}

__global__ void k_item_prob_fast(const double* __restrict__ tsum, int32_t n_clusters, int32_t table_items,
                                 const unsigned long long* __restrict__ counter, double lambda,
                                 double* __restrict__ isum, double* __restrict__ iprob, double* __restrict__ bvec,
                                 double* __restrict__ total_out, unsigned long long* __restrict__ bmin_bits) {
    const int32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    const double total = __ddiv_rn((double)(long long)(*counter), 100.0);   // RM2Job.java:95
    if (i == 0) *total_out = total;
    if (i >= table_items) return;
    double s = 0.0;
    for (int32_t c = 0; c < n_clusters; c++) s += tsum[(size_t)c * table_items + i];
    const double p = (s > 0.0) ? __ddiv_rn(s, total) : 0.0;
    const double b = __dmul_rn(lambda, p);
    isum[i] = s; iprob[i] = p; bvec[i] = b;
    if (s > 0.0) atomicMin(bmin_bits, (unsigned long long)__double_as_longlong(b > 0.0 ? b : 0.0));
}

// number of distinct items of each cluster
__global__ void k_cluster_item_count(const int32_t* __restrict__ tstart, int32_t table_items,
                                     int32_t* __restrict__ icount) {
    const int c = blockIdx.x;
    int cnt = 0;
    for (int32_t i = threadIdx.x; i < table_items; i += blockDim.x) cnt += (tstart[(size_t)c * table_items + i] >= 0);
    __shared__ int s_total;
    if (threadIdx.x == 0) s_total = 0;
    __syncthreads();
    for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    if ((threadIdx.x & 31) == 0) atomicAdd(&s_total, cnt);
    __syncthreads();
    if (threadIdx.x == 0) icount[c] = s_total;
}

// exclusive scan over clusters (tiny) -> item_off; also total
__global__ void k_cluster_offsets(const int32_t* __restrict__ icount, int32_t n_clusters,
                                  int32_t* __restrict__ item_off) {
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        int32_t acc = 0;
        for (int c = 0; c < n_clusters; c++) { item_off[c] = acc; acc += icount[c]; }
        item_off[n_clusters] = acc;
    }
}

// local item numbering per cluster: ascending item id (items[] of createUserAndItemMappings,
// M/rm/AbstractRM2Reducer.java:238-272, canonicalised), plus the per-slot tables.
__global__ void k_local_items(const int32_t* __restrict__ tstart, const int32_t* __restrict__ tend,
                              int32_t table_items, const int32_t* __restrict__ item_off,
                              const double* __restrict__ bvec,
                              int32_t* __restrict__ tloc, int32_t* __restrict__ c_item,
                              int32_t* __restrict__ c_start, int32_t* __restrict__ c_len,
                              double* __restrict__ c_b) {
    const int c = blockIdx.x;
    const int T = blockDim.x;  // 1024
    __shared__ int s_warp[32];
    __shared__ int s_base;
    if (threadIdx.x == 0) s_base = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    for (int32_t i0 = 0; i0 < table_items; i0 += T) {
        const int32_t i = i0 + threadIdx.x;
        const size_t t = (size_t)c * table_items + (i < table_items ? i : 0);
        const int32_t st = (i < table_items) ? tstart[t] : -1;
        const int f = st >= 0;
        const unsigned bal = __ballot_sync(0xffffffffu, f);
        const int inwarp = __popc(bal & ((1u << lane) - 1));
        if (lane == 0) s_warp[w] = __popc(bal);
        __syncthreads();
        int wbase = 0;
        for (int q = 0; q < w; q++) wbase += s_warp[q];
        int tot = 0;
        for (int q = 0; q < (T >> 5); q++) tot += s_warp[q];
        const int base = s_base;
        if (f) {
            const int32_t l = base + wbase + inwarp;
            tloc[t] = l;
            const int32_t slot = item_off[c] + l;
            c_item[slot] = i;
            c_start[slot] = st;
            c_len[slot] = tend[t] - st;
            c_b[slot] = bvec[i];
        }
        __syncthreads();
        if (threadIdx.x == 0) s_base = base + tot;
        __syncthreads();
    }
}

// CSR entry -> local item id, d_vi = P[v][i] - b_i with P rounded exactly as
// probItemGivenUser does: (1 - lambda) * (rating / sum) + lambda * p   (:384-389, no FMA).
__global__ void k_delta(const uint64_t* __restrict__ keys, const float* __restrict__ s_score, int32_t m,
                        int item_bits, const int32_t* __restrict__ rank_cluster,
                        const double* __restrict__ usum, const double* __restrict__ bvec,
                        const int32_t* __restrict__ tloc, int32_t table_items, double lambda,
                        int32_t* __restrict__ csr_loc, double* __restrict__ csr_delta) {
    const int32_t e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= m) return;
    const uint64_t k = keys[e];
    const int32_t item = (int32_t)(k & ((1ull << item_bits) - 1));
    const int32_t rank = (int32_t)(k >> item_bits);
    const int32_t c = rank_cluster[rank];
    const double b = bvec[item];
    const double oml = __dsub_rn(1.0, lambda);
    const double q = __dmul_rn(oml, __ddiv_rn((double)s_score[e], usum[rank]));
    const double P = __dadd_rn(q, b);
    FY_CHECK(item >= 0 && item < table_items && tloc[(size_t)c * table_items + item] >= 0);
    csr_loc[e] = tloc[(size_t)c * table_items + item];
    csr_delta[e] = __dsub_rn(P, b);
}

// CSC side copies: local user index and d for every CSC position
__global__ void k_csc_fill(const uint64_t* __restrict__ keys2, const int32_t* __restrict__ csc_src, int32_t m,
                           int rank_bits, const int32_t* __restrict__ rank_cluster,
                           const int32_t* __restrict__ cstart, const double* __restrict__ csr_delta,
                           int32_t* __restrict__ csc_lu, double* __restrict__ csc_delta) {
    const int32_t x = blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= m) return;
    const int32_t rank = (int32_t)(keys2[x] & ((1ull << rank_bits) - 1));
    csc_lu[x] = rank - cstart[rank_cluster[rank]];
    csc_delta[x] = csr_delta[csc_src[x]];
}

// alpha_j = sum of d over the raters of j, and for every rater u of j
//   c(u,j) = (K-1)*b_j + sum_{v != u, v rated j} d_vj = (K-1)*b_j + (prefix before u + suffix after u)
// one warp per (cluster, item) slot, 32 raters per step with a shuffle scan and a running carry (fixed
// order, so deterministic); all terms non-negative, so no cancellation for any lambda.
__global__ void k_alpha_cuj(const int32_t* __restrict__ c_start, const int32_t* __restrict__ c_len,
                            const double* __restrict__ c_b, const uint64_t* __restrict__ keys2, int rank_bits,
                            const int32_t* __restrict__ rank_cluster, const int32_t* __restrict__ cstart,
                            const int32_t* __restrict__ csc_src, const double* __restrict__ csc_delta,
                            int32_t n_slots, double* __restrict__ c_alpha, double* __restrict__ csr_c,
                            unsigned long long* __restrict__ cbound /* [3*KC]: max alpha, max b, min b (bits) */) {
    const int32_t s = (int32_t)((blockIdx.x * (size_t)blockDim.x + threadIdx.x) >> 5);
    const int lane = threadIdx.x & 31;
    if (s >= n_slots) return;
    const int32_t x0 = c_start[s], n = c_len[s];
    const int32_t c = rank_cluster[(int32_t)(keys2[x0] & ((1ull << rank_bits) - 1))];
    const double km1b = __dmul_rn((double)(cstart[c + 1] - cstart[c] - 1), c_b[s]);
    double carry = 0.0;
    for (int32_t r0 = 0; r0 < n; r0 += 32) {                 // exclusive prefix, ascending
        const int32_t r = r0 + lane;
        const double d = (r < n) ? csc_delta[x0 + r] : 0.0;
        double incl = d;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const double v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl = __dadd_rn(incl, v); }
        double excl = __shfl_up_sync(0xffffffffu, incl, 1);
        if (lane == 0) excl = 0.0;
        if (r < n) csr_c[csc_src[x0 + r]] = __dadd_rn(carry, excl);
        carry = __dadd_rn(carry, __shfl_sync(0xffffffffu, incl, 31));
    }
    if (lane == 0) {
        c_alpha[s] = carry;
        // per-cluster bounds used to pick the scale / peel period of the float score kernel
        const unsigned long long bb = (unsigned long long)__double_as_longlong(c_b[s]);
        atomicMax(&cbound[3 * c], (unsigned long long)__double_as_longlong(carry));
        atomicMax(&cbound[3 * c + 1], bb);
        atomicMin(&cbound[3 * c + 2], bb);
    }
    carry = 0.0;
    for (int32_t r0 = 0; r0 < n; r0 += 32) {                 // exclusive suffix, descending
        const int32_t r = n - 1 - (r0 + lane);
        const double d = (r >= 0) ? csc_delta[x0 + r] : 0.0;
        double incl = d;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const double v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl = __dadd_rn(incl, v); }
        double excl = __shfl_up_sync(0xffffffffu, incl, 1);
        if (lane == 0) excl = 0.0;
        if (r >= 0) {
            const int32_t e = csc_src[x0 + r];
            csr_c[e] = __dadd_rn(km1b, __dadd_rn(csr_c[e], __dadd_rn(carry, excl)));
        }
        carry = __dadd_rn(carry, __shfl_sync(0xffffffffu, incl, 31));
    }
}

// sort key (cluster, most raters first) of every (cluster, item) slot -> processing order of the H-build rows: a popular
// row walks thousands of raters one after the other and must not be the last one to start
__global__ void k_row_perm_keys(const int32_t* __restrict__ c_len, const int32_t* __restrict__ item_off, int32_t n_clusters,
                                int32_t n_slots, uint64_t* __restrict__ keys, int32_t* __restrict__ vals) {
    const int32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n_slots) return;
    int32_t lo = 0, hi = n_clusters;                         // cluster of the slot: last c with item_off[c] <= s
    while (hi - lo > 1) { const int32_t mid = (lo + hi) >> 1; if (item_off[mid] <= s) lo = mid; else hi = mid; }
    keys[s] = ((uint64_t)(uint32_t)lo << 32) | (uint64_t)(0xffffffffu - (uint32_t)c_len[s]);
    vals[s] = s;
}

// sort key (cluster, most rated items first) of every user rank -> processing order of the score kernel
__global__ void k_perm_keys(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ rank_cluster, int32_t n_users,
                            uint64_t* __restrict__ keys, int32_t* __restrict__ vals) {
    const int32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_users) return;
    const uint32_t n = (uint32_t)(rowptr[r + 1] - rowptr[r]);
    keys[r] = ((uint64_t)(uint32_t)rank_cluster[r] << 32) | (uint64_t)(0xffffffffu - n);
    vals[r] = r;
}

// per-user work estimate n_u * I_c (for sharding) as double
__global__ void k_user_work(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ rank_cluster,
                            const int32_t* __restrict__ icount, int32_t n_users, double* __restrict__ work) {
    const int32_t u = blockIdx.x * blockDim.x + threadIdx.x;
    if (u >= n_users) return;
    work[u] = (double)(rowptr[u + 1] - rowptr[u]) * (double)icount[rank_cluster[u]];
}

// ---------------------------------------------------------------------------------------------
// H[j][i] = S[j][i] + b_j * alpha_i   for one cluster.  One CTA per (row j, column chunk).
// The row accumulates in shared memory; raters of j are visited in ascending user order and each
// rater's own rated items (sorted) are scattered by distinct threads, so the sum order -- hence
// every bit of H -- is fixed, and two identical item columns give identical H columns (exact ties
// stay exact ties, as in the reference's double loop).
// ---------------------------------------------------------------------------------------------
// upper 32 bits of a positive double, rounded to nearest on the dropped half: 20 mantissa bits,
// relative error <= 2^-21.  The hi-word plane of H is what the approximate score kernel streams.
__device__ __forceinline__ uint32_t hi_word_rn(double x) {
    return (uint32_t)(((unsigned long long)__double_as_longlong(x) + 0x80000000ull) >> 32);
}

// first CSR entry of every (user, column slice) of one cluster; slice q covers local item ids
// [q*slice_w, (q+1)*slice_w); n_bound = number of slices + 1
__global__ void k_chunk_ptr(int32_t rank0, int32_t K_c, int32_t n_bound, int32_t slice_w,
                            const int32_t* __restrict__ rowptr, const int32_t* __restrict__ csr_loc,
                            int32_t* __restrict__ chunk_ptr) {
    const int32_t idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= K_c * n_bound) return;
    const int32_t u = idx / n_bound, q = idx % n_bound;
    int32_t a = rowptr[rank0 + u], b = rowptr[rank0 + u + 1];
    const int32_t target = q * slice_w;
    while (a < b) { const int32_t mid = (a + b) >> 1; if (csr_loc[mid] < target) a = mid + 1; else b = mid; }
    chunk_ptr[idx] = a;
}

// One WARP builds one (row j, 512-column slice) of H; a CTA is just 8 such independent tasks, so
// there is no block barrier anywhere and ~56 warps per SM hide each other's latency.
// The warp walks the flattened list of (rater of j, that rater's entries inside the slice) 32
// entries at a time: the raters of j are taken 32 at a time, lane r holding rater r's delta and
// entry range; a 5-step shuffle binary search maps a flat position to its rater.  Two lanes can
// meet on one column only for different raters; they are then added in lane order (= ascending
// rater), so every accumulator's summation order is fixed.
__global__ void __launch_bounds__(H_THREADS)
k_build_H(int32_t I_c, int32_t ld, int32_t n_slices, int32_t slot0,
          const int32_t* __restrict__ c_start, const int32_t* __restrict__ c_len,
          const double* __restrict__ c_b, const double* __restrict__ c_alpha,
          const int32_t* __restrict__ csc_lu, const double* __restrict__ csc_delta,
          const int32_t* __restrict__ chunk_ptr, const int32_t* __restrict__ csr_loc,
          const double* __restrict__ csr_delta, double* __restrict__ H, uint32_t* __restrict__ Hh,
          int plane_mode /* 1 = hi words, 2 = float(H * plane_scale) */, double plane_scale) {
    extern __shared__ double acc_all[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int32_t j = blockIdx.x;
    const int32_t sl = blockIdx.y * H_WARPS + warp;
    if (sl >= n_slices) return;
    double* __restrict__ acc = acc_all + warp * H_SLICE;
    const int32_t c0 = sl * H_SLICE;
    const int32_t w = min(H_SLICE, I_c - c0);
    const int32_t n_bound = n_slices + 1;
#pragma unroll
    for (int t = lane * 2; t < H_SLICE; t += 64) *reinterpret_cast<double2*>(acc + t) = make_double2(0.0, 0.0);
    __syncwarp();
    const int32_t x0 = c_start[slot0 + j], nr = c_len[slot0 + j];
    for (int32_t g = 0; g < nr; g += 32) {
        // lane r: rater g+r of row j
        int32_t lo = 0, cnt = 0;
        double d = 0.0;
        if (g + lane < nr) {
            const int32_t x = x0 + g + lane;
            const size_t bp = (size_t)csc_lu[x] * n_bound + sl;
            d = csc_delta[x];
            lo = chunk_ptr[bp];
            cnt = chunk_ptr[bp + 1] - lo;
        }
        int32_t incl = cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int32_t v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
        const int32_t off = incl - cnt;                                   // exclusive prefix
        const int32_t total = __shfl_sync(0xffffffffu, incl, 31);
        const int32_t ebase = lo - off;                                   // entry = ebase[r] + f
        for (int32_t f0 = 0; f0 < total; f0 += 32) {
            const int32_t f = f0 + lane;
            const bool valid = f < total;
            int r = 0;                                                    // largest r with off[r] <= f
#pragma unroll
            for (int st = 16; st > 0; st >>= 1) {
                const int32_t v = __shfl_sync(0xffffffffu, off, r + st);
                if (v <= f) r += st;
            }
            const int32_t e = __shfl_sync(0xffffffffu, ebase, r) + f;
            const double dr = __shfl_sync(0xffffffffu, d, r);
            int32_t col = -1 - lane;                                      // unique dummy for idle lanes
            double val = 0.0;
            if (valid) {
                col = csr_loc[e] - c0;
                FY_CHECK(col >= 0 && col < w);
                val = __dmul_rn(dr, csr_delta[e]);
            }
            const unsigned peers = __match_any_sync(0xffffffffu, col);
            if (valid) {
                if (peers == (1u << lane)) {
                    acc[col] = __dadd_rn(acc[col], val);
                } else {                                                  // several raters, one column
                    const int leader = __ffs(peers) - 1;
                    double sum = (lane == leader) ? acc[col] : 0.0;
                    for (unsigned mm = peers; mm; mm &= mm - 1) {
                        const double v = __shfl_sync(peers, val, __ffs(mm) - 1);
                        sum = __dadd_rn(sum, v);
                    }
                    if (lane == leader) acc[col] = sum;
                }
            }
            __syncwarp();
        }
    }
    const double bj = c_b[slot0 + j];
    double* __restrict__ row = H + (size_t)j * ld + c0;
    const double* __restrict__ al = c_alpha + slot0 + c0;
    uint32_t* __restrict__ rowh = Hh ? Hh + (size_t)j * ld + c0 : nullptr;
    for (int t = lane * 2; t < w; t += 64) {
        if (t + 1 < w) {
            double2 o;
            o.x = __fma_rn(bj, al[t], acc[t]);
            o.y = __fma_rn(bj, al[t + 1], acc[t + 1]);
            *reinterpret_cast<double2*>(row + t) = o;
            if (rowh) *reinterpret_cast<uint2*>(rowh + t) = (plane_mode == 2)
                ? make_uint2(__float_as_uint((float)(o.x * plane_scale)), __float_as_uint((float)(o.y * plane_scale)))
                : make_uint2(hi_word_rn(o.x), hi_word_rn(o.y));
        } else {
            const double o = __fma_rn(bj, al[t], acc[t]);
            row[t] = o;
            if (rowh) rowh[t] = (plane_mode == 2) ? __float_as_uint((float)(o * plane_scale)) : hi_word_rn(o);
        }
    }
    if (sl == n_slices - 1)
        for (int32_t i = I_c + lane; i < ld; i += 32) {                              // padding columns
            H[(size_t)j * ld + i] = 1.0;
            if (Hh) Hh[(size_t)j * ld + i] = (plane_mode == 2) ? 0x3f800000u : 0x3ff00000u;
        }
}

// ---------------------------------------------------------------------------------------------
// k_build_H2 (round 2): the same H, bit for bit, without the flattened-list search.
// One warp owns (row j, a range of RW columns) with an fp64 accumulator row in shared memory and walks the
// raters of j strictly in ascending order, ONE rater at a time, lanes over that rater's entries inside the
// range (a rater's columns are distinct: no conflict, no match_any, no binary search, no shuffle scan).
// The first 32 entries of the NEXT rater are gathered into registers before the current rater is
// accumulated, so one L2 gather is always in flight per warp; ~28 warps per SM hide the rest.
// Per rater the warp issues ~30 instructions (4 shuffles, 2 gathers, 1 multiply, 1 shared-memory
// read-add-write) against ~100 per 32 flattened entries in k_build_H; the price is lane use (the
// work-weighted rater has ~15 entries in a 1024-column range at ML-20M shape, ~60 at Netflix shape).
// The write-out handles 4 columns per lane per step and is specialised on the plane mode.
// Summation order per accumulator = ascending rater, separate multiply and add: identical to k_build_H.
// ---------------------------------------------------------------------------------------------
// Two light raters per step (lanes 0-15 / 16-31, one match_any to order a shared column): bit-identical, measured SLOWER
// (2.92 vs 2.64 ms per ML-20M-sized cluster, 5.73 vs 4.85 ms Netflix-sized) -- the kernel is not bound by its step count.
// Compiled out; -DFY_H2_PAIR=1 brings the experiment back.
#ifndef FY_H2_PAIR
#define FY_H2_PAIR 0
#endif
#ifndef FY_H2_FLOAT_DIRECT
#define FY_H2_FLOAT_DIRECT 1
#endif
constexpr bool H2_FLOAT_DIRECT = FY_H2_FLOAT_DIRECT != 0;   // bulk write-out: 4-byte plane by ordinary stores in the finishing pass
constexpr bool H2_PAIR = FY_H2_PAIR != 0;

template <int RW, int NW, int PM /* 0 = fp64 plane only, 1 = + hi words, 2 = + float(H * plane_scale) */, bool BULK /* write-out by cp.async.bulk */>
__global__ void __launch_bounds__(NW * 32)
k_build_H2(int32_t I_c, int32_t ld, int32_t n_ranges, int32_t slot0,
           const int32_t* __restrict__ c_start, const int32_t* __restrict__ c_len,
           const double* __restrict__ c_b, const double* __restrict__ c_alpha,
           const int32_t* __restrict__ csc_lu, const double* __restrict__ csc_delta,
           const int32_t* __restrict__ chunk_ptr, const int32_t* __restrict__ csr_loc,
           const double* __restrict__ csr_delta, double* __restrict__ H, uint32_t* __restrict__ Hh,
           double plane_scale, int32_t nchunk /* CTAs per row; the grid is linear, row-major: the CTAs in flight write neighbouring memory */,
           const int32_t* __restrict__ row_perm /* slots of the cluster, most raters first (null = item order) */) {
    static_assert(RW % 128 == 0, "the write-out takes 128 columns per warp step");
    extern __shared__ __align__(16) unsigned char h2_smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int32_t j, qb;
    if (nchunk > 0) { j = (int32_t)(blockIdx.x / (unsigned)nchunk); qb = (int32_t)(blockIdx.x % (unsigned)nchunk); }
    else { j = blockIdx.x; qb = blockIdx.y; }
    if (row_perm) j = row_perm[slot0 + j] - slot0;
    const int32_t q = qb * NW + warp;
    if (q >= n_ranges) return;
    double* __restrict__ acc = reinterpret_cast<double*>(h2_smem) + (size_t)warp * RW;
    const int32_t c0 = q * RW;
    const int32_t w = min(RW, I_c - c0);
    const int32_t nb = n_ranges + 1;
    const int32_t x0 = c_start[slot0 + j], nr = c_len[slot0 + j];

    // lane r of a group: rater g + r of row j -> its delta and its entry range inside this column range
    int32_t lo_n = 0, cnt_n = 0;
    double d_n = 0.0;
    if (lane < nr) {
        const int32_t x = x0 + lane;
        const size_t bp = (size_t)csc_lu[x] * nb + q;
        d_n = csc_delta[x];
        lo_n = chunk_ptr[bp];
        cnt_n = chunk_ptr[bp + 1] - lo_n;
    }
#pragma unroll 4
    for (int t = lane * 2; t < RW; t += 64) *reinterpret_cast<double2*>(acc + t) = make_double2(0.0, 0.0);
    __syncwarp();

    const double* __restrict__ accb = acc - c0;                          // indexed by the cluster-local column
    double* __restrict__ accw = acc - c0;
    for (int32_t g = 0; g < nr; g += 32) {
        const int32_t lo_l = lo_n, cnt_l = cnt_n;
        const double d_l = d_n;
        lo_n = 0; cnt_n = 0; d_n = 0.0;
        if (g + 32 + lane < nr) {                                        // next group's registers, in flight during this group
            const int32_t x = x0 + g + 32 + lane;
            const size_t bp = (size_t)csc_lu[x] * nb + q;
            d_n = csc_delta[x];
            lo_n = chunk_ptr[bp];
            cnt_n = chunk_ptr[bp + 1] - lo_n;
        }
        unsigned m = __ballot_sync(0xffffffffu, cnt_l > 0);              // raters with entries here; bit order = ascending rater
        if (m == 0u) continue;
        // Raters are taken in batches of PF; the gathers of batch B are issued before batch A is accumulated and
        // vice versa, so PF..2*PF gathers are in flight per warp (the single-step look-ahead was latency bound:
        // 25 % of the stall samples sat on the first use of the gathered column).
        constexpr int PF = 4;
        int32_t loA[PF], cntA[PF], cn2A[PF], colA[PF], loB[PF], cntB[PF], cn2B[PF], colB[PF];
        double dA[PF], dlA[PF], dB[PF], dlB[PF];
        // A slot holds one rater (lanes = its first 32 entries) or, when two consecutive raters have <= 16 entries each in this
        // range (the common case at ML-20M shape: 12 of 32 lanes were busy per step), BOTH: lanes 0-15 the earlier rater,
        // lanes 16-31 the later one (cn2 > 0).  The earlier rater's update of a column both touch is applied first.
        auto fetch = [&](int32_t (&lo)[PF], int32_t (&cnt)[PF], int32_t (&cn2)[PF], int32_t (&col)[PF], double (&d)[PF], double (&dl)[PF]) {
#pragma unroll
            for (int z = 0; z < PF; z++) {
                cnt[z] = 0; cn2[z] = 0; lo[z] = 0; col[z] = c0; d[z] = 0.0; dl[z] = 0.0;
                if (m) {
                    const int r = __ffs(m) - 1;
                    m &= m - 1;
                    const int32_t lo1 = __shfl_sync(0xffffffffu, lo_l, r);
                    const int32_t c1 = __shfl_sync(0xffffffffu, cnt_l, r);
                    const double d1 = __shfl_sync(0xffffffffu, d_l, r);
                    int32_t lo2 = 0, c2 = 0;
                    double d2 = 0.0;
                    if (H2_PAIR && c1 <= 16 && m) {
                        const int r2 = __ffs(m) - 1;
                        const int32_t cc = __shfl_sync(0xffffffffu, cnt_l, r2);
                        if (cc <= 16) {
                            m &= m - 1;
                            c2 = cc;
                            lo2 = __shfl_sync(0xffffffffu, lo_l, r2);
                            d2 = __shfl_sync(0xffffffffu, d_l, r2);
                        }
                    }
                    lo[z] = lo1; cnt[z] = c1; cn2[z] = c2;
                    if (c2 > 0) {
                        const bool hi = lane >= 16;
                        const int l16 = lane & 15;
                        const int32_t mylo = hi ? lo2 : lo1, mycnt = hi ? c2 : c1;
                        d[z] = hi ? d2 : d1;
                        if (l16 < mycnt) { col[z] = csr_loc[mylo + l16]; dl[z] = csr_delta[mylo + l16]; }
                    } else {
                        d[z] = d1;
                        if (lane < c1) { col[z] = csr_loc[lo1 + lane]; dl[z] = csr_delta[lo1 + lane]; }
                    }
                }
            }
        };
        auto consume = [&](const int32_t (&lo)[PF], const int32_t (&cnt)[PF], const int32_t (&cn2)[PF], const int32_t (&col)[PF],
                           const double (&d)[PF], const double (&dl)[PF]) {
#pragma unroll
            for (int z = 0; z < PF; z++) {
                if (cn2[z] > 0) {                                        // two raters in one step
                    const bool valid = (lane & 15) < ((lane >= 16) ? cn2[z] : cnt[z]);
                    FY_CHECK(!valid || (col[z] >= c0 && col[z] < c0 + w));
                    const unsigned peers = __match_any_sync(0xffffffffu, valid ? col[z] : (-1 - lane));
                    const bool first = (peers & ((1u << lane) - 1u)) == 0u;  // the earlier rater's lane (or the only one) of this column
                    if (valid && first) accw[col[z]] = __dadd_rn(accb[col[z]], __dmul_rn(d[z], dl[z]));
                    __syncwarp();
                    if (__any_sync(0xffffffffu, valid && !first)) {
                        if (valid && !first) accw[col[z]] = __dadd_rn(accb[col[z]], __dmul_rn(d[z], dl[z]));
                        __syncwarp();
                    }
                } else if (cnt[z] > 0) {
                    FY_CHECK(lane >= cnt[z] || (col[z] >= c0 && col[z] < c0 + w));
                    if (lane < cnt[z]) accw[col[z]] = __dadd_rn(accb[col[z]], __dmul_rn(d[z], dl[z]));
                    if (cnt[z] > 32) {                                   // a heavy rater: its further entries, 4 gathers in flight
                        for (int32_t k = 32 + lane; k < cnt[z]; k += 128) {
                            int32_t c2[4];
                            double v2[4];
#pragma unroll
                            for (int y = 0; y < 4; y++) {
                                const bool in = k + 32 * y < cnt[z];
                                c2[y] = in ? csr_loc[lo[z] + k + 32 * y] : -1;
                                v2[y] = in ? csr_delta[lo[z] + k + 32 * y] : 0.0;
                            }
#pragma unroll
                            for (int y = 0; y < 4; y++)
                                if (c2[y] >= 0) { FY_CHECK(c2[y] >= c0 && c2[y] < c0 + w); accw[c2[y]] = __dadd_rn(accb[c2[y]], __dmul_rn(d[z], v2[y])); }
                        }
                    }
                    __syncwarp();                                        // the next rater may hit columns other lanes just updated
                }
            }
        };
        fetch(loA, cntA, cn2A, colA, dA, dlA);
        for (;;) {
            fetch(loB, cntB, cn2B, colB, dB, dlB);
            consume(loA, cntA, cn2A, colA, dA, dlA);
            if (cntB[0] == 0) break;
            fetch(loA, cntA, cn2A, colA, dA, dlA);
            consume(loB, cntB, cn2B, colB, dB, dlB);
            if (cntA[0] == 0) break;
        }
    }
    __syncwarp();
    // write-out: H = acc + b_j * alpha (one fma, as k_build_H), 4 columns per lane per step
    const double bj = c_b[slot0 + j];
    double* __restrict__ row = H + (size_t)j * ld + c0;
    const double* __restrict__ al = c_alpha + slot0 + c0;
    uint32_t* __restrict__ rowh = (PM != 0) ? Hh + (size_t)j * ld + c0 : nullptr;
    if (BULK) {
        // The row segment is finished in place in shared memory and leaves through the bulk-copy engine: 8.4 GB of
        // 16-byte LSU stores per cluster queued in the same L1 FIFO as every gather of the SM (ncu: the loads of the
        // write-out and of the accumulate phase waited thousands of cycles behind them).
        const int32_t wp = min(RW, ld - c0);                             // padding columns of the last range included
#pragma unroll 4
        for (int t = lane * 2; t < wp; t += 64) {
            double2 a = *reinterpret_cast<const double2*>(acc + t);
            a.x = (t < w) ? __fma_rn(bj, __ldg(al + t), a.x) : 1.0;
            a.y = (t + 1 < w) ? __fma_rn(bj, __ldg(al + t + 1), a.y) : 1.0;
            *reinterpret_cast<double2*>(acc + t) = a;
            if (PM != 0 && H2_FLOAT_DIRECT) {
                // the 4-byte plane leaves in the same pass with ordinary 8-byte stores (a third of the bytes): converting it in
                // place for a second bulk copy costs another read and write of the accumulator row in shared memory
                uint2 o = (PM == 2)
                    ? make_uint2(__float_as_uint((float)(a.x * plane_scale)), __float_as_uint((float)(a.y * plane_scale)))
                    : make_uint2(hi_word_rn(a.x), hi_word_rn(a.y));
                if (PM == 2 && t >= w) o = make_uint2(0x3f800000u, 0x3f800000u);     // padding columns: 1.0f, as k_build_H
                *reinterpret_cast<uint2*>(rowh + t) = o;
            }
        }
        asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
        __syncwarp();
        if (lane == 0) {
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;\n"
                         :: "l"(row), "r"((unsigned)__cvta_generic_to_shared(acc)), "r"((unsigned)wp * 8u) : "memory");
            asm volatile("cp.async.bulk.commit_group;\n" ::: "memory");
        }
        if (PM != 0 && !H2_FLOAT_DIRECT) {
            // 4-byte plane: converted in place, front to back (step s reads bytes [512 s, 512 s + 512) and writes
            // [256 s, 256 s + 256), which every earlier step has already consumed), after the fp64 copy has read the row
            if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;\n" ::: "memory");
            __syncwarp();
            uint32_t* __restrict__ accf = reinterpret_cast<uint32_t*>(acc);
            for (int t = lane * 2; t < wp; t += 64) {
                const double2 a = *reinterpret_cast<const double2*>(acc + t);
                __syncwarp();
                uint2 o = (PM == 2)
                    ? make_uint2(__float_as_uint((float)(a.x * plane_scale)), __float_as_uint((float)(a.y * plane_scale)))
                    : make_uint2(hi_word_rn(a.x), hi_word_rn(a.y));
                if (PM == 2 && t >= w) o = make_uint2(0x3f800000u, 0x3f800000u);     // padding columns: 1.0f, as k_build_H
                *reinterpret_cast<uint2*>(accf + t) = o;
            }
            asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
            __syncwarp();
            if (lane == 0) {
                asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;\n"
                             :: "l"(rowh), "r"((unsigned)__cvta_generic_to_shared(acc)), "r"((unsigned)wp * 4u) : "memory");
                asm volatile("cp.async.bulk.commit_group;\n" ::: "memory");
            }
        }
        if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;\n" ::: "memory");   // shared memory is released at exit
        __syncwarp();
        return;
    }
    const int32_t w4 = w & ~3;
#pragma unroll 4
    for (int t = lane * 4; t < w4; t += 128) {
        const double a0 = __ldg(al + t), a1 = __ldg(al + t + 1), a2 = __ldg(al + t + 2), a3 = __ldg(al + t + 3);
        const double2 s0 = *reinterpret_cast<const double2*>(acc + t);
        const double2 s1 = *reinterpret_cast<const double2*>(acc + t + 2);
        double2 o0, o1;
        o0.x = __fma_rn(bj, a0, s0.x); o0.y = __fma_rn(bj, a1, s0.y);
        o1.x = __fma_rn(bj, a2, s1.x); o1.y = __fma_rn(bj, a3, s1.y);
        *reinterpret_cast<double2*>(row + t) = o0;
        *reinterpret_cast<double2*>(row + t + 2) = o1;
        if (PM == 2)
            *reinterpret_cast<uint4*>(rowh + t) = make_uint4(__float_as_uint((float)(o0.x * plane_scale)), __float_as_uint((float)(o0.y * plane_scale)),
                                                              __float_as_uint((float)(o1.x * plane_scale)), __float_as_uint((float)(o1.y * plane_scale)));
        else if (PM == 1)
            *reinterpret_cast<uint4*>(rowh + t) = make_uint4(hi_word_rn(o0.x), hi_word_rn(o0.y), hi_word_rn(o1.x), hi_word_rn(o1.y));
    }
    if (lane < w - w4) {                                                 // last 1..3 columns of the last range
        const int t = w4 + lane;
        const double o = __fma_rn(bj, al[t], acc[t]);
        row[t] = o;
        if (PM == 2) rowh[t] = __float_as_uint((float)(o * plane_scale));
        else if (PM == 1) rowh[t] = hi_word_rn(o);
    }
    if (q == n_ranges - 1)
        for (int32_t i = I_c + lane; i < ld; i += 32) {                              // padding columns
            H[(size_t)j * ld + i] = 1.0;
            if (PM == 2) Hh[(size_t)j * ld + i] = 0x3f800000u;
            else if (PM == 1) Hh[(size_t)j * ld + i] = 0x3ff00000u;
        }
}

// ---------------------------------------------------------------------------------------------
// Score kernel: CTA = (user, tile of 256 candidates).  Streams H[j][tile] for every rated j.
//   prod_i *= H[j][i] + b_i * c(u,j);   score = log(prod) + pvpi
// The product is kept as mantissa * 2^ex; the exponent is peeled every L factors (L is chosen by
// the host from a lower bound on t so that L factors can neither overflow nor underflow), so one
// log per (u,i) replaces n_u logs (SURVEY.md 7.3 "log throughput").
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint64_t desc_key(double s) {
    const uint64_t b = (uint64_t)__double_as_longlong(s);
    const uint64_t asc = (b >> 63) ? ~b : (b | 0x8000000000000000ull);
    return ~asc;   // smaller key = larger score
}
__device__ __forceinline__ double key_to_score(uint64_t key) {
    const uint64_t asc = ~key;
    const uint64_t b = (asc >> 63) ? (asc & 0x7fffffffffffffffull) : ~asc;
    return __longlong_as_double((long long)b);
}

__device__ __forceinline__ void peel_exponent(double& p, int& ex) {
    const int hi = __double2hiint(p);
    ex += (hi >> 20) - 1023;
    p = __hiloint2double((hi & 0x000fffff) | 0x3ff00000, __double2loint(p));
}

__device__ __forceinline__ double2 ld_row(const double* p) {
    return __ldg(reinterpret_cast<const double2*>(p));
}

template <int L>
__global__ void __launch_bounds__(SCORE_THREADS)
k_score(const double* __restrict__ H, int32_t I_c, int32_t ld, int32_t rank_begin, int32_t slot0,
        const int32_t* __restrict__ rowptr, const int32_t* __restrict__ csr_loc,
        const double* __restrict__ csr_c, const double* __restrict__ c_b,
        double log_items, double log_K, double* __restrict__ scores, unsigned long long* __restrict__ ustat) {
    __shared__ int32_t s_j[SCORE_CHUNK];
    __shared__ double s_c[SCORE_CHUNK];
    __shared__ unsigned s_rated[SCORE_TILE / 32];

    const int32_t rank = rank_begin + blockIdx.x;
    const int32_t tile0 = blockIdx.y * SCORE_TILE;
    const int32_t i = tile0 + 2 * threadIdx.x;
    const int32_t e0 = rowptr[rank];
    const int32_t n = rowptr[rank + 1] - e0;

    if (threadIdx.x < SCORE_TILE / 32) s_rated[threadIdx.x] = 0u;
    const double b0 = (i < I_c) ? c_b[slot0 + i] : 0.0;
    const double b1 = (i + 1 < I_c) ? c_b[slot0 + i + 1] : 0.0;
    double p0 = 1.0, p1 = 1.0;
    int ex0 = 0, ex1 = 0;
    bool z0 = false, z1 = false;
    const double* __restrict__ Hc = H + i;

    for (int32_t base = 0; base < n; base += SCORE_CHUNK) {
        const int32_t cnt = min(SCORE_CHUNK, n - base);
        __syncthreads();
        for (int32_t k = threadIdx.x; k < cnt; k += SCORE_THREADS) {
            const int32_t j = csr_loc[e0 + base + k];
            s_j[k] = j;
            s_c[k] = csr_c[e0 + base + k];
            const int32_t d = j - tile0;
            if (d >= 0 && d < SCORE_TILE) atomicOr(&s_rated[d >> 5], 1u << (d & 31));
        }
        __syncthreads();
        int32_t k = 0;
        for (; k + 8 <= cnt; k += 8) {
            double2 h[8];
#pragma unroll
            for (int q = 0; q < 8; q++) h[q] = ld_row(Hc + (size_t)s_j[k + q] * ld);
#pragma unroll
            for (int q = 0; q < 8; q++) {
                const double c = s_c[k + q];
                p0 *= fma(b0, c, h[q].x);
                p1 *= fma(b1, c, h[q].y);
                if ((q + 1) % L == 0) {
                    if (L == 1) { if (p0 == 0.0) { z0 = true; p0 = 1.0; } if (p1 == 0.0) { z1 = true; p1 = 1.0; } }
                    peel_exponent(p0, ex0);
                    peel_exponent(p1, ex1);
                }
            }
        }
        for (; k < cnt; k++) {
            const double2 h = ld_row(Hc + (size_t)s_j[k] * ld);
            const double c = s_c[k];
            p0 *= fma(b0, c, h.x);
            p1 *= fma(b1, c, h.y);
            if (L == 1) { if (p0 == 0.0) { z0 = true; p0 = 1.0; } if (p1 == 0.0) { z1 = true; p1 = 1.0; } }
            if (L == 1 || ((k & 7) + 1) % L == 0 || k + 1 == cnt) { peel_exponent(p0, ex0); peel_exponent(p1, ex1); }
        }
    }
    __syncthreads();
    // pvpi = (n - 1) * log(numberOfItems) - n * log(K)      (AbstractRM2Reducer.java:328-329)
    const double pvpi = __dsub_rn(__dmul_rn((double)(n - 1), log_items), __dmul_rn((double)n, log_K));
    const double LN2_HI = 6.93147180369123816490e-01, LN2_LO = 1.90821492927058770002e-10;
    const double NANV = __longlong_as_double(0x7ff8000000000000ll);
    const double NINF = __longlong_as_double(0xfff0000000000000ll);
    double s0 = fma((double)ex0, LN2_HI, fma((double)ex0, LN2_LO, log(p0))) + pvpi;
    double s1 = fma((double)ex1, LN2_HI, fma((double)ex1, LN2_LO, log(p1))) + pvpi;
    if (z0) s0 = NINF;
    if (z1) s1 = NINF;
    const int d = 2 * threadIdx.x;
    const unsigned word = s_rated[d >> 5];
    if ((word >> (d & 31)) & 1u) s0 = NANV;
    if ((word >> ((d + 1) & 31)) & 1u) s1 = NANV;
    if (i >= I_c) s0 = NANV;
    if (i + 1 >= I_c) s1 = NANV;
    double2 out; out.x = s0; out.y = s1;
    *reinterpret_cast<double2*>(scores + (size_t)blockIdx.x * ld + i) = out;
    // per-user candidate count and key range for the radix select (saves k_topn one pass)
    const bool v0 = (s0 == s0), v1 = (s1 == s1);
    unsigned long long kmin = ~0ull, kmax = 0ull;
    if (v0) { const unsigned long long k = desc_key(s0); kmin = k; kmax = k; }
    if (v1) { const unsigned long long k = desc_key(s1); kmin = min(kmin, k); kmax = max(kmax, k); }
    int cnt = (int)v0 + (int)v1;
    for (int o = 16; o > 0; o >>= 1) {
        cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
        kmin = min(kmin, __shfl_xor_sync(0xffffffffu, kmin, o));
        kmax = max(kmax, __shfl_xor_sync(0xffffffffu, kmax, o));
    }
    if ((threadIdx.x & 31) == 0 && cnt > 0) {
        unsigned long long* st = ustat + 3 * (size_t)blockIdx.x;
        atomicAdd(st, (unsigned long long)cnt);
        atomicMin(st + 1, kmin);
        atomicMax(st + 2, kmax);
    }
}

// ---------------------------------------------------------------------------------------------
// Approximate score kernel: identical to k_score but streams the hi-word plane of H (4 bytes per
// log-term, 20-bit mantissa): CTA = (user, 512 candidates), 16-byte loads of 4 candidates.
// |log t~ - log t| <= 2^-21 per term, so |score~ - score| <= n_u * 2^-21 (+ fp64 noise): a rigorous
// bound that k_topn's margin pass turns into a candidate set containing the exact top-N; k_refine then
// re-scores those candidates from the fp64 plane.  Only used when every t > 0 (L >= 2).
// ---------------------------------------------------------------------------------------------
template <int L>
__global__ void __launch_bounds__(SCORE_THREADS)
k_score_hi(const uint32_t* __restrict__ Hh, int32_t I_c, int32_t ld, int32_t rank_begin, int32_t slot0,
           const int32_t* __restrict__ rowptr, const int32_t* __restrict__ csr_loc,
           const double* __restrict__ csr_c, const double* __restrict__ c_b,
           double log_items, double log_K, double* __restrict__ scores, unsigned long long* __restrict__ ustat) {
    __shared__ int32_t s_j[SCORE_CHUNK];
    __shared__ double s_c[SCORE_CHUNK];
    __shared__ unsigned s_rated[SCOREH_TILE / 32];

    const int32_t rank = rank_begin + blockIdx.x;
    const int32_t tile0 = blockIdx.y * SCOREH_TILE;
    const int32_t i = tile0 + 4 * threadIdx.x;
    const int32_t e0 = rowptr[rank];
    const int32_t n = rowptr[rank + 1] - e0;

    if (threadIdx.x < SCOREH_TILE / 32) s_rated[threadIdx.x] = 0u;
    double b[4], p[4];
    int ex[4];
#pragma unroll
    for (int q = 0; q < 4; q++) { b[q] = (i + q < I_c) ? c_b[slot0 + i + q] : 0.0; p[q] = 1.0; ex[q] = 0; }
    const uint32_t* __restrict__ Hc = Hh + i;

    for (int32_t base = 0; base < n; base += SCORE_CHUNK) {
        const int32_t cnt = min(SCORE_CHUNK, n - base);
        __syncthreads();
        for (int32_t k = threadIdx.x; k < cnt; k += SCORE_THREADS) {
            const int32_t j = csr_loc[e0 + base + k];
            s_j[k] = j;
            s_c[k] = csr_c[e0 + base + k];
            const int32_t d = j - tile0;
            if (d >= 0 && d < SCOREH_TILE) atomicOr(&s_rated[d >> 5], 1u << (d & 31));
        }
        __syncthreads();
        int32_t k = 0;
        for (; k + 8 <= cnt; k += 8) {
            uint4 h[8];
#pragma unroll
            for (int q = 0; q < 8; q++) h[q] = __ldg(reinterpret_cast<const uint4*>(Hc + (size_t)s_j[k + q] * ld));
#pragma unroll
            for (int q = 0; q < 8; q++) {
                const double c = s_c[k + q];
                p[0] *= fma(b[0], c, __hiloint2double((int)h[q].x, 0));
                p[1] *= fma(b[1], c, __hiloint2double((int)h[q].y, 0));
                p[2] *= fma(b[2], c, __hiloint2double((int)h[q].z, 0));
                p[3] *= fma(b[3], c, __hiloint2double((int)h[q].w, 0));
                if ((q + 1) % L == 0) {
#pragma unroll
                    for (int z = 0; z < 4; z++) peel_exponent(p[z], ex[z]);
                }
            }
        }
        for (; k < cnt; k++) {
            const uint4 h = __ldg(reinterpret_cast<const uint4*>(Hc + (size_t)s_j[k] * ld));
            const double c = s_c[k];
            p[0] *= fma(b[0], c, __hiloint2double((int)h.x, 0));
            p[1] *= fma(b[1], c, __hiloint2double((int)h.y, 0));
            p[2] *= fma(b[2], c, __hiloint2double((int)h.z, 0));
            p[3] *= fma(b[3], c, __hiloint2double((int)h.w, 0));
            if (((k & 7) + 1) % L == 0 || k + 1 == cnt) {
#pragma unroll
                for (int z = 0; z < 4; z++) peel_exponent(p[z], ex[z]);
            }
        }
    }
    __syncthreads();
    const double pvpi = __dsub_rn(__dmul_rn((double)(n - 1), log_items), __dmul_rn((double)n, log_K));
    const double LN2_HI = 6.93147180369123816490e-01, LN2_LO = 1.90821492927058770002e-10;
    const double NANV = __longlong_as_double(0x7ff8000000000000ll);
    double s[4];
    const int d = 4 * threadIdx.x;
    const unsigned word = s_rated[d >> 5];
    unsigned long long kmin = ~0ull, kmax = 0ull;
    int cnt = 0;
#pragma unroll
    for (int q = 0; q < 4; q++) {
        s[q] = fma((double)ex[q], LN2_HI, fma((double)ex[q], LN2_LO, log(p[q]))) + pvpi;
        if (((word >> ((d + q) & 31)) & 1u) || i + q >= I_c) s[q] = NANV;
        if (s[q] == s[q]) { const unsigned long long k = desc_key(s[q]); kmin = min(kmin, k); kmax = max(kmax, k); cnt++; }
    }
    double* dst = scores + (size_t)blockIdx.x * ld + i;
    *reinterpret_cast<double2*>(dst) = make_double2(s[0], s[1]);
    *reinterpret_cast<double2*>(dst + 2) = make_double2(s[2], s[3]);
    for (int o = 16; o > 0; o >>= 1) {
        cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
        kmin = min(kmin, __shfl_xor_sync(0xffffffffu, kmin, o));
        kmax = max(kmax, __shfl_xor_sync(0xffffffffu, kmax, o));
    }
    if ((threadIdx.x & 31) == 0 && cnt > 0) {
        unsigned long long* st = ustat + 3 * (size_t)blockIdx.x;
        atomicAdd(st, (unsigned long long)cnt);
        atomicMin(st + 1, kmin);
        atomicMax(st + 2, kmax);
    }
}

// Float variant of the approximate score kernel: the 4-byte plane holds float(H * 2^s) (s chosen per
// cluster so that t * 2^s is centred on 1), t' = fmaf(b_i 2^s, c_uj, h') and the running product are
// fp32 (full-rate pipe, no 64-bit register pairs to assemble), the exponent is peeled every LF factors.
// Every factor carries <= 4 * 2^-24 relative error: |score~ - score| <= n_u * 2.4e-7.
__device__ __forceinline__ void peel_exponent_f(float& p, int& ex) {
    const int b = __float_as_int(p);
    ex += (b >> 23) - 127;
    p = __int_as_float((b & 0x007fffff) | 0x3f800000);
}

// sm_100 packed fp32: one FFMA2 / FMUL2 does two candidates (each lane rounds exactly like the scalar FFMA / FMUL), which
// halves the arithmetic issue slots of a kernel that the plane-read probe shows to be issue bound, not L2 bound
typedef unsigned long long f32x2_t;
__device__ __forceinline__ f32x2_t pack_f32x2(float lo, float hi) {
    return ((f32x2_t)__float_as_uint(hi) << 32) | (f32x2_t)__float_as_uint(lo);
}
__device__ __forceinline__ f32x2_t pack_u32x2(uint32_t lo, uint32_t hi) { return ((f32x2_t)hi << 32) | (f32x2_t)lo; }
__device__ __forceinline__ f32x2_t ffma2(f32x2_t a, f32x2_t b, f32x2_t c) {
    f32x2_t r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}
__device__ __forceinline__ f32x2_t fmul2(f32x2_t a, f32x2_t b) {
    f32x2_t r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ void peel_exponent_f2(f32x2_t& p, int& e0, int& e1) {
    uint32_t lo = (uint32_t)p, hi = (uint32_t)(p >> 32);
    e0 += (int)(lo >> 23) - 127;                        // p > 0: no sign bit
    e1 += (int)(hi >> 23) - 127;
    lo = (lo & 0x007fffffu) | 0x3f800000u;
    hi = (hi & 0x007fffffu) | 0x3f800000u;
    p = pack_u32x2(lo, hi);
}

template <int LF, int MINB = 1>
__global__ void __launch_bounds__(SCORE_THREADS, MINB)
k_score_f32(const uint32_t* __restrict__ Hf, int32_t I_c, int32_t ld, int32_t rank_begin, int32_t slot0,
            const int32_t* __restrict__ rowptr, const int32_t* __restrict__ csr_loc,
            const double* __restrict__ csr_c, const double* __restrict__ c_b, double plane_scale, int32_t scale_exp,
            double log_items, double log_K, double* __restrict__ scores, unsigned long long* __restrict__ ustat,
            const int32_t* __restrict__ perm /* users of the batch, most active first (null = rank order) */) {
    __shared__ uint2 s_jc[SCORE_CHUNK];                  // (local row j, float bits of c(u,j)): one 8-byte load per row
    __shared__ unsigned s_rated[SCOREH_TILE / 32];

    // longest-processing-time-first: the CTA of a user with thousands of rated items must not start last
    const int32_t bx = perm ? perm[rank_begin + blockIdx.x] - rank_begin : (int32_t)blockIdx.x;
    const int32_t rank = rank_begin + bx;
    const int32_t tile0 = blockIdx.y * SCOREH_TILE;
    const int32_t i = tile0 + 4 * threadIdx.x;
    const int32_t e0 = rowptr[rank];
    const int32_t n = rowptr[rank + 1] - e0;

    if (threadIdx.x < SCOREH_TILE / 32) s_rated[threadIdx.x] = 0u;
    float b[4];
    int ex[4];
#pragma unroll
    for (int q = 0; q < 4; q++) { b[q] = (i + q < I_c) ? (float)(c_b[slot0 + i + q] * plane_scale) : 0.0f; ex[q] = 0; }
    const f32x2_t b01 = pack_f32x2(b[0], b[1]), b23 = pack_f32x2(b[2], b[3]);
    f32x2_t p01 = pack_f32x2(1.0f, 1.0f), p23 = p01;
    const uint32_t* __restrict__ Hc = Hf + i;

    for (int32_t base = 0; base < n; base += SCORE_CHUNK) {
        const int32_t cnt = min(SCORE_CHUNK, n - base);
        __syncthreads();
        for (int32_t k = threadIdx.x; k < cnt; k += SCORE_THREADS) {
            const int32_t j = csr_loc[e0 + base + k];
            FY_CHECK(j >= 0 && j < I_c && k < SCORE_CHUNK);
            s_jc[k] = make_uint2((uint32_t)j, __float_as_uint((float)csr_c[e0 + base + k]));
            const int32_t d = j - tile0;
            if (d >= 0 && d < SCOREH_TILE) atomicOr(&s_rated[d >> 5], 1u << (d & 31));
        }
        __syncthreads();
        int32_t k = 0;
        for (; k + 8 <= cnt; k += 8) {
            uint4 h[8];
            uint32_t cb[8];
#pragma unroll
            for (int q = 0; q < 8; q++) {
                const uint2 jc = s_jc[k + q];
                cb[q] = jc.y;
                h[q] = __ldg(reinterpret_cast<const uint4*>(Hc + (size_t)jc.x * ld));
            }
#pragma unroll
            for (int q = 0; q < 8; q++) {
                const f32x2_t c2 = pack_u32x2(cb[q], cb[q]);
                p01 = fmul2(p01, ffma2(b01, c2, pack_u32x2(h[q].x, h[q].y)));     // p *= fmaf(b, c, h), two candidates per instruction
                p23 = fmul2(p23, ffma2(b23, c2, pack_u32x2(h[q].z, h[q].w)));
                if ((q + 1) % LF == 0) { peel_exponent_f2(p01, ex[0], ex[1]); peel_exponent_f2(p23, ex[2], ex[3]); }
            }
        }
        for (; k < cnt; k++) {
            const uint2 jc = s_jc[k];
            const uint4 h = __ldg(reinterpret_cast<const uint4*>(Hc + (size_t)jc.x * ld));
            const f32x2_t c2 = pack_u32x2(jc.y, jc.y);
            p01 = fmul2(p01, ffma2(b01, c2, pack_u32x2(h.x, h.y)));
            p23 = fmul2(p23, ffma2(b23, c2, pack_u32x2(h.z, h.w)));
            if (((k & 7) + 1) % LF == 0 || k + 1 == cnt) { peel_exponent_f2(p01, ex[0], ex[1]); peel_exponent_f2(p23, ex[2], ex[3]); }
        }
    }
    float p[4];
    p[0] = __uint_as_float((uint32_t)p01); p[1] = __uint_as_float((uint32_t)(p01 >> 32));
    p[2] = __uint_as_float((uint32_t)p23); p[3] = __uint_as_float((uint32_t)(p23 >> 32));
    __syncthreads();
    const double pvpi = __dsub_rn(__dmul_rn((double)(n - 1), log_items), __dmul_rn((double)n, log_K));
    const double LN2_HI = 6.93147180369123816490e-01, LN2_LO = 1.90821492927058770002e-10;
    const double NANV = __longlong_as_double(0x7ff8000000000000ll);
    double s[4];
    const int d = 4 * threadIdx.x;
    const unsigned word = s_rated[d >> 5];
    unsigned long long kmin = ~0ull, kmax = 0ull;
    int cnt = 0;
#pragma unroll
    for (int q = 0; q < 4; q++) {
        const double et = (double)(ex[q] - n * scale_exp);               // undo the 2^s on each of the n factors
        s[q] = fma(et, LN2_HI, fma(et, LN2_LO, log((double)p[q]))) + pvpi;
        if (((word >> ((d + q) & 31)) & 1u) || i + q >= I_c) s[q] = NANV;
        if (s[q] == s[q]) { const unsigned long long k = desc_key(s[q]); kmin = min(kmin, k); kmax = max(kmax, k); cnt++; }
    }
    double* dst = scores + (size_t)bx * ld + i;
    *reinterpret_cast<double2*>(dst) = make_double2(s[0], s[1]);
    *reinterpret_cast<double2*>(dst + 2) = make_double2(s[2], s[3]);
    for (int o = 16; o > 0; o >>= 1) {
        cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
        kmin = min(kmin, __shfl_xor_sync(0xffffffffu, kmin, o));
        kmax = max(kmax, __shfl_xor_sync(0xffffffffu, kmax, o));
    }
    if ((threadIdx.x & 31) == 0 && cnt > 0) {
        unsigned long long* st = ustat + 3 * (size_t)bx;
        atomicAdd(st, (unsigned long long)cnt);
        atomicMin(st + 1, kmin);
        atomicMax(st + 2, kmax);
    }
}

// ---------------------------------------------------------------------------------------------
// k_score_f32_tma (round 2): the same arithmetic as k_score_f32, with the plane rows brought in by the bulk-copy engine.
// The plane-read probe showed that the 16-byte __ldg version is latency bound, not L2 bound (3.4 ms against 2.1 ms for the
// same loads without arithmetic): a warp alternates "issue 8 loads - wait - 110 instructions" and the other resident warps
// do not cover the waits.  Here a row segment (512 candidates x 4 B = 2 KB, contiguous) is ONE cp.async.bulk into a
// shared-memory ring of SCORE_TMA_STAGES stages x 8 rows, issued by a dedicated producer warp and tracked by full / empty
// mbarriers; the four consumer warps only read shared memory (LDS.128) and do arithmetic, so the copies of the next two
// stages are always in flight while a stage is consumed, and the per-row global-load and address instructions disappear.
// MEASURED (one B200): slower than the __ldg kernel, 3.83 vs 3.38 ms per ML-20M-sized cluster and 11.1 vs 8.0 ms per
// Netflix-sized one.  No byte of a row segment is shared between threads, so the ring is a second trip through shared
// memory: 2 x 43 GB per cluster at 128 B/clk/SM is 2.4 ms before any arithmetic.  Bulk staging pays where a tile is
// re-read by many threads (the GEMM operands), not for a stream that every thread reads once.  Kept behind
// FY_SCORE_TMA=1 as the record of that experiment; the product path is k_score_f32.
// ---------------------------------------------------------------------------------------------
constexpr int SCORE_TMA_STAGES = 3;
constexpr int SCORE_TMA_ROWS = 8;                      // rows per stage
constexpr int SCORE_TMA_THREADS = SCORE_THREADS + 32;  // 4 consumer warps + 1 producer warp
constexpr size_t SCORE_TMA_SMEM = (size_t)SCORE_TMA_STAGES * SCORE_TMA_ROWS * SCOREH_TILE * 4;

__device__ __forceinline__ void mbar_init_s(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_s(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx_s(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait_s(uint32_t bar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile("{\n\t.reg .pred p;\n\t"
                     "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                     "selp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    } while (!done);
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

template <int LF>
__global__ void __launch_bounds__(SCORE_TMA_THREADS)
k_score_f32_tma(const uint32_t* __restrict__ Hf, int32_t I_c, int32_t ld, int32_t rank_begin, int32_t slot0,
                const int32_t* __restrict__ rowptr, const int32_t* __restrict__ csr_loc,
                const double* __restrict__ csr_c, const double* __restrict__ c_b, double plane_scale, int32_t scale_exp,
                double log_items, double log_K, double* __restrict__ scores, unsigned long long* __restrict__ ustat,
                const int32_t* __restrict__ perm /* users of the batch, most active first (null = rank order) */) {
    extern __shared__ __align__(128) unsigned char tma_ring[];           // [stages][rows][512] uint32
    __shared__ uint2 s_jc[SCORE_CHUNK];                                  // (local row j, float bits of c(u,j))
    __shared__ unsigned s_rated[SCOREH_TILE / 32];
    __shared__ __align__(8) unsigned long long s_bar[2 * SCORE_TMA_STAGES];

    const int32_t bx = perm ? perm[rank_begin + blockIdx.x] - rank_begin : (int32_t)blockIdx.x;
    const int32_t rank = rank_begin + bx;
    const int32_t tile0 = blockIdx.y * SCOREH_TILE;
    const int32_t e0 = rowptr[rank];
    const int32_t n = rowptr[rank + 1] - e0;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const bool producer = (warp == SCORE_THREADS / 32);
    const uint32_t ring = (uint32_t)__cvta_generic_to_shared(tma_ring);
    const uint32_t bars = (uint32_t)__cvta_generic_to_shared(s_bar);
    auto full = [&](int st) { return bars + 8u * (uint32_t)st; };
    auto empty = [&](int st) { return bars + 8u * (uint32_t)(SCORE_TMA_STAGES + st); };

    if (threadIdx.x == 0) {
        for (int st = 0; st < SCORE_TMA_STAGES; st++) { mbar_init_s(full(st), 1); mbar_init_s(empty(st), SCORE_THREADS / 32); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (threadIdx.x < SCOREH_TILE / 32) s_rated[threadIdx.x] = 0u;

    const int32_t i = tile0 + 4 * (int32_t)threadIdx.x;                  // consumers only (threadIdx.x < 128)
    float b[4] = {0.0f, 0.0f, 0.0f, 0.0f};
    int ex[4] = {0, 0, 0, 0};
    if (!producer) {
#pragma unroll
        for (int q = 0; q < 4; q++) b[q] = (i + q < I_c) ? (float)(c_b[slot0 + i + q] * plane_scale) : 0.0f;
    }
    const f32x2_t b01 = pack_f32x2(b[0], b[1]), b23 = pack_f32x2(b[2], b[3]);
    f32x2_t p01 = pack_f32x2(1.0f, 1.0f), p23 = p01;
    const uint32_t* __restrict__ Hrow = Hf + tile0;                      // + j * ld: the 2 KB segment of row j
    uint32_t step = 0;                                                   // stage counter, runs across chunks (ring position / phase)

    for (int32_t base = 0; base < n; base += SCORE_CHUNK) {
        const int32_t cnt = min(SCORE_CHUNK, n - base);
        __syncthreads();                                                 // everyone is done with the previous chunk's list
        if (!producer) {
            for (int32_t k = threadIdx.x; k < cnt; k += SCORE_THREADS) {
                const int32_t j = csr_loc[e0 + base + k];
                FY_CHECK(j >= 0 && j < I_c && k < SCORE_CHUNK);
                s_jc[k] = make_uint2((uint32_t)j, __float_as_uint((float)csr_c[e0 + base + k]));
                const int32_t d = j - tile0;
                if (d >= 0 && d < SCOREH_TILE) atomicOr(&s_rated[d >> 5], 1u << (d & 31));
            }
        }
        __syncthreads();
        const int32_t n_stage = (cnt + SCORE_TMA_ROWS - 1) / SCORE_TMA_ROWS;
        if (producer) {
            // lanes 0..7 each copy one row of the stage; lane 0 registers the stage's byte count first
            for (int32_t sg = 0; sg < n_stage; sg++) {
                const uint32_t it = step + (uint32_t)sg;
                const int st = (int)(it % SCORE_TMA_STAGES);
                const uint32_t ph = (it / SCORE_TMA_STAGES) & 1u;
                const int32_t rows = min(SCORE_TMA_ROWS, cnt - sg * SCORE_TMA_ROWS);
                if (lane == 0) {
                    mbar_wait_s(empty(st), ph ^ 1u);                     // the four consumer warps have left this stage
                    mbar_expect_tx_s(full(st), (uint32_t)rows * SCOREH_TILE * 4u);
                }
                __syncwarp();
                if (lane < rows) {
                    const uint32_t j = s_jc[sg * SCORE_TMA_ROWS + lane].x;
                    bulk_g2s(ring + (uint32_t)((st * SCORE_TMA_ROWS + lane) * SCOREH_TILE * 4), Hrow + (size_t)j * ld, SCOREH_TILE * 4u, full(st));
                }
            }
        } else {
            for (int32_t sg = 0; sg < n_stage; sg++) {
                const uint32_t it = step + (uint32_t)sg;
                const int st = (int)(it % SCORE_TMA_STAGES);
                const uint32_t ph = (it / SCORE_TMA_STAGES) & 1u;
                const int32_t rows = min(SCORE_TMA_ROWS, cnt - sg * SCORE_TMA_ROWS);
                mbar_wait_s(full(st), ph);
                const uint4* __restrict__ rowbase = reinterpret_cast<const uint4*>(tma_ring) + (size_t)(st * SCORE_TMA_ROWS) * (SCOREH_TILE / 4) + threadIdx.x;
                const uint2* __restrict__ jc = s_jc + sg * SCORE_TMA_ROWS;
                if (rows == SCORE_TMA_ROWS) {
#pragma unroll
                    for (int q = 0; q < SCORE_TMA_ROWS; q++) {
                        const uint4 h = rowbase[q * (SCOREH_TILE / 4)];
                        const uint32_t cb = jc[q].y;
                        const f32x2_t c2 = pack_u32x2(cb, cb);
                        p01 = fmul2(p01, ffma2(b01, c2, pack_u32x2(h.x, h.y)));
                        p23 = fmul2(p23, ffma2(b23, c2, pack_u32x2(h.z, h.w)));
                        if ((q + 1) % LF == 0) { peel_exponent_f2(p01, ex[0], ex[1]); peel_exponent_f2(p23, ex[2], ex[3]); }
                    }
                } else {
                    for (int q = 0; q < rows; q++) {
                        const uint4 h = rowbase[q * (SCOREH_TILE / 4)];
                        const uint32_t cb = jc[q].y;
                        const f32x2_t c2 = pack_u32x2(cb, cb);
                        p01 = fmul2(p01, ffma2(b01, c2, pack_u32x2(h.x, h.y)));
                        p23 = fmul2(p23, ffma2(b23, c2, pack_u32x2(h.z, h.w)));
                        if ((q + 1) % LF == 0 || q + 1 == rows) { peel_exponent_f2(p01, ex[0], ex[1]); peel_exponent_f2(p23, ex[2], ex[3]); }
                    }
                }
                __syncwarp();
                if (lane == 0) mbar_arrive_s(empty(st));                 // this warp has read the stage
            }
        }
        step += (uint32_t)n_stage;
    }
    __syncthreads();
    if (producer) return;
    float p[4];
    p[0] = __uint_as_float((uint32_t)p01); p[1] = __uint_as_float((uint32_t)(p01 >> 32));
    p[2] = __uint_as_float((uint32_t)p23); p[3] = __uint_as_float((uint32_t)(p23 >> 32));
    const double pvpi = __dsub_rn(__dmul_rn((double)(n - 1), log_items), __dmul_rn((double)n, log_K));
    const double LN2_HI = 6.93147180369123816490e-01, LN2_LO = 1.90821492927058770002e-10;
    const double NANV = __longlong_as_double(0x7ff8000000000000ll);
    double s[4];
    const int d = 4 * threadIdx.x;
    const unsigned word = s_rated[d >> 5];
    unsigned long long kmin = ~0ull, kmax = 0ull;
    int cnt = 0;
#pragma unroll
    for (int q = 0; q < 4; q++) {
        const double et = (double)(ex[q] - n * scale_exp);               // undo the 2^s on each of the n factors
        s[q] = fma(et, LN2_HI, fma(et, LN2_LO, log((double)p[q]))) + pvpi;
        if (((word >> ((d + q) & 31)) & 1u) || i + q >= I_c) s[q] = NANV;
        if (s[q] == s[q]) { const unsigned long long k = desc_key(s[q]); kmin = min(kmin, k); kmax = max(kmax, k); cnt++; }
    }
    double* dst = scores + (size_t)bx * ld + i;
    *reinterpret_cast<double2*>(dst) = make_double2(s[0], s[1]);
    *reinterpret_cast<double2*>(dst + 2) = make_double2(s[2], s[3]);
    for (int o = 16; o > 0; o >>= 1) {
        cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
        kmin = min(kmin, __shfl_xor_sync(0xffffffffu, kmin, o));
        kmax = max(kmax, __shfl_xor_sync(0xffffffffu, kmax, o));
    }
    if ((threadIdx.x & 31) == 0 && cnt > 0) {
        unsigned long long* st = ustat + 3 * (size_t)bx;
        atomicAdd(st, (unsigned long long)cnt);
        atomicMin(st + 1, kmin);
        atomicMax(st + 2, kmax);
    }
}

// exact fp64 re-score of the candidates: grid (user, candidate group), one warp per candidate, lanes
// stride the rated items (4 gathers in flight per lane); the lane products are combined by a fixed
// butterfly, so the result is deterministic and identical item columns give identical scores.
__global__ void __launch_bounds__(REFINE_THREADS)
k_refine_score(const double* __restrict__ H, int32_t ld, int32_t rank_begin, int32_t slot0,
               const int32_t* __restrict__ rowptr, const int32_t* __restrict__ csr_loc, const double* __restrict__ csr_c,
               const double* __restrict__ c_b, double log_items, double log_K,
               int32_t cap, const int32_t* __restrict__ cand, const int32_t* __restrict__ cand_cnt,
               double* __restrict__ cand_score) {
    const int32_t cnt = cand_cnt[blockIdx.x];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int32_t q = blockIdx.y * (REFINE_THREADS / 32) + warp;
    if (cnt > cap || q >= cnt) return;
    const int32_t rank = rank_begin + blockIdx.x;
    const int32_t e0 = rowptr[rank], n = rowptr[rank + 1] - e0;
    const int32_t i = cand[(size_t)blockIdx.x * cap + q];
    FY_CHECK(i >= 0 && q < cap);
    const double bi = c_b[slot0 + i];
    const double* __restrict__ Hi = H + i;
    double p = 1.0;
    int ex = 0;
    int32_t k = lane;
    for (; k + 96 < n; k += 128) {
        double h[4], c[4];
#pragma unroll
        for (int z = 0; z < 4; z++) { h[z] = Hi[(size_t)csr_loc[e0 + k + 32 * z] * ld]; c[z] = csr_c[e0 + k + 32 * z]; }
#pragma unroll
        for (int z = 0; z < 4; z++) { p *= fma(bi, c[z], h[z]); peel_exponent(p, ex); }
    }
    for (; k < n; k += 32) { p *= fma(bi, csr_c[e0 + k], Hi[(size_t)csr_loc[e0 + k] * ld]); peel_exponent(p, ex); }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double po = __shfl_xor_sync(0xffffffffu, p, o);
        const int eo = __shfl_xor_sync(0xffffffffu, ex, o);
        p *= po; ex += eo;
        peel_exponent(p, ex);
    }
    if (lane == 0) {
        const double pvpi = __dsub_rn(__dmul_rn((double)(n - 1), log_items), __dmul_rn((double)n, log_K));
        const double LN2_HI = 6.93147180369123816490e-01, LN2_LO = 1.90821492927058770002e-10;
        cand_score[(size_t)blockIdx.x * cap + q] = fma((double)ex, LN2_HI, fma((double)ex, LN2_LO, log(p))) + pvpi;
    }
}

// final (score desc, item id asc) order of the re-scored candidates -> the user's top-N
__global__ void __launch_bounds__(REFINE_THREADS)
k_refine_sort(int32_t rank_begin, int32_t ub, int32_t slot0, const int32_t* __restrict__ c_item,
              int32_t cap, const int32_t* __restrict__ cand, const int32_t* __restrict__ cand_cnt,
              const double* __restrict__ cand_score, int32_t out_stride, int32_t* __restrict__ out_item,
              double* __restrict__ out_score, const int32_t* __restrict__ out_count) {
    extern __shared__ unsigned char smem_raw[];
    const int32_t orow = rank_begin + blockIdx.x - ub;
    const int32_t n_out = out_count[orow];
    const int32_t cnt = cand_cnt[blockIdx.x];
    if (n_out == 0 || cnt > cap) return;                           // nothing to do / overflow (run is redone exactly)
    int P2 = 1; while (P2 < cnt) P2 <<= 1;
    uint64_t* sk = reinterpret_cast<uint64_t*>(smem_raw);
    int32_t* si = reinterpret_cast<int32_t*>(smem_raw + (size_t)P2 * sizeof(uint64_t));
    const int tid = threadIdx.x;
    for (int t = tid; t < P2; t += REFINE_THREADS) {
        if (t < cnt) { sk[t] = desc_key(cand_score[(size_t)blockIdx.x * cap + t]); si[t] = cand[(size_t)blockIdx.x * cap + t]; }
        else { sk[t] = ~0ull; si[t] = 0x7fffffff; }
    }
    __syncthreads();
    for (int k = 2; k <= P2; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int t = tid; t < P2; t += REFINE_THREADS) {
                const int ixj = t ^ j;
                if (ixj > t) {
                    const uint64_t ka = sk[t], kb = sk[ixj];
                    const int32_t ia = si[t], ib = si[ixj];
                    const bool a_gt_b = (ka > kb) || (ka == kb && ia > ib);
                    const bool up = ((t & k) == 0);
                    if (a_gt_b == up) { sk[t] = kb; sk[ixj] = ka; si[t] = ib; si[ixj] = ia; }
                }
            }
            __syncthreads();
        }
    }
    for (int t = tid; t < n_out; t += REFINE_THREADS) {
        FY_CHECK(t < P2 && t < out_stride && si[t] != 0x7fffffff);
        out_item[(size_t)orow * out_stride + t] = c_item[slot0 + si[t]];
        out_score[(size_t)orow * out_stride + t] = key_to_score(sk[t]);
    }
}

// ---------------------------------------------------------------------------------------------
// Top-N per user: radix select on the order-preserving 64-bit image of the score, then a bitonic
// sort of the selected (key, index) pairs.  Order = (score desc, item id asc); local index order
// is item id order.  NaN marks "not a candidate" (rated by the user / padding).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(TOPN_THREADS)
k_topn(const double* __restrict__ scores, const unsigned long long* __restrict__ ustat,
       int32_t I_c, int32_t ld, int32_t rank_begin, int32_t slot0,
       int32_t top_n, int32_t out_stride, int32_t filter_users, int32_t split, int32_t n_splits,
       const int32_t* __restrict__ rank_userid, const int32_t* __restrict__ c_item,
       int32_t out_row0, int32_t* __restrict__ out_item, double* __restrict__ out_score,
       int32_t* __restrict__ out_count,
       // optional (auto mode): the candidates that can be in the EXACT top-N, i.e. approximate score >= (N-th approximate
       // score) - 2 eps, collected in one more pass over the row while it is still in L2 (was a kernel of its own)
       const int32_t* __restrict__ rowptr, int32_t cap, double eps_per_term, int32_t* __restrict__ cand,
       int32_t* __restrict__ cand_cnt, int* __restrict__ overflow) {
    extern __shared__ unsigned char smem_raw[];
    constexpr int RBITS = 11, NBINS = 1 << RBITS;             // 4 bins per thread
    static_assert(NBINS == 4 * TOPN_THREADS, "bucket search assumes 4 bins per thread");
    __shared__ unsigned s_hist[NBINS];
    __shared__ int s_wsum[TOPN_THREADS / 32];
    __shared__ int s_digit, s_remaining, s_done, s_nsel, s_eqtaken;

    const int32_t rank = rank_begin + blockIdx.x;
    const int32_t orow = out_row0 + blockIdx.x;
    const double* __restrict__ row = scores + (size_t)blockIdx.x * ld;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;

    const int32_t uid = rank_userid[rank];
    const int c_u = (int)ustat[3 * (size_t)blockIdx.x];
    const int n_out = min(top_n, c_u);
    // split filter :203-205, "no unrated item" :210-213, filterUsers :220-223
    if (uid < filter_users || (n_splits > 1 && (uid % n_splits) != split) || n_out == 0) {
        if (tid == 0) { out_count[orow] = 0; if (cand_cnt) cand_cnt[blockIdx.x] = 0; }
        return;
    }
    const uint64_t kbase = ustat[3 * (size_t)blockIdx.x + 1];
    const uint64_t span = ustat[3 * (size_t)blockIdx.x + 2] - kbase;
    // radix select on d = key - kbase, RBITS bits per pass from the highest set bit of the span
    int shift = (span == 0) ? 0 : (64 - __clzll((long long)span));   // bits below `shift` are undecided
    uint64_t prefix = 0;                                             // decided value of d >> shift
    const bool all = (n_out == c_u);
    if (tid == 0) { s_remaining = n_out; s_done = all ? 1 : 0; }
    __syncthreads();
    while (!s_done && shift > 0) {
        const int nshift = max(shift - RBITS, 0);
        const int width = shift - nshift;
        for (int t = tid; t < NBINS; t += TOPN_THREADS) s_hist[t] = 0;
        __syncthreads();
        for (int32_t i = tid; i < I_c; i += TOPN_THREADS) {
            const double sc = row[i];
            if (sc == sc) {
                const uint64_t d = desc_key(sc) - kbase;
                if ((shift >= 64 ? 0ull : (d >> shift)) == prefix)
                    atomicAdd(&s_hist[(unsigned)((d >> nshift) & ((1u << width) - 1))], 1u);
            }
        }
        __syncthreads();
        // block-wide search of the bucket holding the s_remaining-th smallest element
        const int rem = s_remaining;
        const int h0 = (int)s_hist[4 * tid], h1 = (int)s_hist[4 * tid + 1], h2 = (int)s_hist[4 * tid + 2], h3 = (int)s_hist[4 * tid + 3];
        const int mine = h0 + h1 + h2 + h3;
        int incl = mine;
        for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
        if (lane == 31) s_wsum[wid] = incl;
        __syncthreads();
        int excl = incl - mine;
        for (int q = 0; q < wid; q++) excl += s_wsum[q];
        if (excl < rem && rem <= excl + mine) {          // exactly one thread
            const int hh[4] = {h0, h1, h2, h3};
            int acc = excl;
            for (int q = 0; q < 4; q++) {
                if (acc + hh[q] >= rem) {
                    s_digit = 4 * tid + q;
                    s_remaining = rem - acc;                      // how many to take from this bucket
                    if (hh[q] == rem - acc) s_done = 1;          // the whole bucket is taken
                    break;
                }
                acc += hh[q];
            }
        }
        __syncthreads();
        prefix = (prefix << width) | (uint64_t)s_digit;
        shift = nshift;
        __syncthreads();
    }
    // Selection rule with q = (key - kbase) >> shift:
    //   q < prefix -> selected;  q == prefix -> selected if the whole bucket is taken, else the
    //   first s_remaining of them in index order (only reachable with shift == 0: equal keys, so
    //   index order = ascending item id is the canonical tie-break).
    const bool whole_bucket = (s_done != 0);
    const int eq_take = s_remaining;

    int P2 = 1; while (P2 < n_out) P2 <<= 1;
    uint64_t* sk = reinterpret_cast<uint64_t*>(smem_raw);
    int32_t* si = reinterpret_cast<int32_t*>(smem_raw + (size_t)P2 * sizeof(uint64_t));
    for (int t = tid; t < P2; t += TOPN_THREADS) { sk[t] = ~0ull; si[t] = 0x7fffffff; }
    if (tid == 0) { s_nsel = 0; s_eqtaken = 0; }
    __syncthreads();
    for (int32_t i0 = 0; i0 < I_c; i0 += TOPN_THREADS) {
        const int32_t i = i0 + tid;
        bool take = false, eq = false;
        uint64_t key = 0;
        if (i < I_c) {
            const double sc = row[i];
            if (sc == sc) {
                key = desc_key(sc);
                if (all) take = true;
                else {
                    const uint64_t q = (shift >= 64) ? 0ull : ((key - kbase) >> shift);
                    if (q < prefix) take = true;
                    else if (q == prefix) { if (whole_bucket) take = true; else eq = true; }
                }
            }
        }
        if (!all && !whole_bucket) {
            // ordered take of the first eq_take "equal" elements
            const unsigned bal = __ballot_sync(0xffffffffu, eq);
            if (lane == 0) s_wsum[wid] = __popc(bal);
            __syncthreads();
            int before = s_eqtaken, tot = 0;
            for (int q = 0; q < TOPN_THREADS / 32; q++) { const int v = s_wsum[q]; if (q < wid) before += v; tot += v; }
            if (eq && before + __popc(bal & ((1u << lane) - 1)) < eq_take) take = true;
            __syncthreads();
            if (tid == 0) s_eqtaken += tot;
            __syncthreads();
        }
        if (take) {
            const int pos = atomicAdd(&s_nsel, 1);
            FY_CHECK(pos < P2 && i < I_c);
            if (pos < P2) { sk[pos] = key; si[pos] = i; }
        }
    }
    __syncthreads();
    // bitonic sort by (key, index)
    for (int k = 2; k <= P2; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int t = tid; t < P2; t += TOPN_THREADS) {
                const int ixj = t ^ j;
                if (ixj > t) {
                    const uint64_t ka = sk[t], kb = sk[ixj];
                    const int32_t ia = si[t], ib = si[ixj];
                    const bool a_gt_b = (ka > kb) || (ka == kb && ia > ib);
                    const bool up = ((t & k) == 0);
                    if (a_gt_b == up) { sk[t] = kb; sk[ixj] = ka; si[t] = ib; si[ixj] = ia; }
                }
            }
            __syncthreads();
        }
    }
    for (int t = tid; t < n_out; t += TOPN_THREADS) {
        FY_CHECK(si[t] >= 0 && si[t] < I_c && t < out_stride);
        out_item[(size_t)orow * out_stride + t] = c_item[slot0 + si[t]];
        out_score[(size_t)orow * out_stride + t] = key_to_score(sk[t]);
    }
    if (tid == 0) out_count[orow] = n_out;
    if (cand) {
        const double n_u = (double)(rowptr[rank + 1] - rowptr[rank]);
        const double eps = n_u * eps_per_term + 1e-9;              // rigorous per-term bound + fp64 slack
        const double thr = key_to_score(sk[n_out - 1]) - 2.0 * eps;
        __syncthreads();                                           // everyone has read sk[n_out - 1]; s_nsel is free
        if (tid == 0) s_nsel = 0;
        __syncthreads();
        for (int32_t i = tid; i < I_c; i += TOPN_THREADS) {
            const double sc = row[i];
            if (sc == sc && sc >= thr) {
                const int pos = atomicAdd(&s_nsel, 1);
                if (pos < cap) cand[(size_t)blockIdx.x * cap + pos] = i;
            }
        }
        __syncthreads();
        if (tid == 0) {
            cand_cnt[blockIdx.x] = s_nsel;
            if (s_nsel > cap) atomicAdd(overflow, 1);
        }
    }
}

// ---- large-N path (min(N, I_c) > TOPN_MAX_SELECT): whole-row stable segmented sort ----
__global__ void k_sort_prepare(const double* __restrict__ scores, int32_t I_c, int32_t ld, int64_t n,
                               uint64_t* __restrict__ keys, int32_t* __restrict__ idx) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    const int32_t i = (int32_t)(t % ld);
    const double s = scores[t];
    keys[t] = (i < I_c && s == s) ? desc_key(s) : ~0ull;      // non-candidates sort last
    idx[t] = i;
}

// first min(N, c_u) entries of every user's sorted row (stable sort: equal scores keep ascending item id)
__global__ void k_emit_sorted(const uint64_t* __restrict__ keys, const int32_t* __restrict__ idx,
                              const unsigned long long* __restrict__ ustat, int32_t ld, int32_t rank_begin, int32_t ub,
                              int32_t slot0, int32_t top_n, int32_t out_stride, int32_t filter_users, int32_t split,
                              int32_t n_splits, const int32_t* __restrict__ rank_userid, const int32_t* __restrict__ c_item,
                              int32_t* __restrict__ out_item, double* __restrict__ out_score, int32_t* __restrict__ out_count) {
    const int32_t rank = rank_begin + blockIdx.x, orow = rank - ub;
    const int32_t uid = rank_userid[rank];
    const int c_u = (int)ustat[3 * (size_t)blockIdx.x];
    int n_out = min(top_n, c_u);
    if (uid < filter_users || (n_splits > 1 && (uid % n_splits) != split)) n_out = 0;
    for (int32_t t = threadIdx.x; t < n_out; t += blockDim.x) {
        out_item[(size_t)orow * out_stride + t] = c_item[slot0 + idx[(size_t)blockIdx.x * ld + t]];
        out_score[(size_t)orow * out_stride + t] = key_to_score(keys[(size_t)blockIdx.x * ld + t]);
    }
    if (threadIdx.x == 0) out_count[orow] = n_out;
}

__global__ void k_seg_offsets(int32_t* __restrict__ off, int32_t n, int32_t ld) {
    const int32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i <= n) off[i] = i * ld;
}

__global__ void k_init_ustat(unsigned long long* __restrict__ ustat, int32_t n) {
    const int32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { ustat[3 * (size_t)i] = 0ull; ustat[3 * (size_t)i + 1] = ~0ull; ustat[3 * (size_t)i + 2] = 0ull; }
}

// ---------------------------------------------------------------------------------------------
// Config 3 helpers (the GEMM itself is in cooc_tcgen05.cu)
// ---------------------------------------------------------------------------------------------
// binarised, transposed rating matrix Bt[item][user] (uint8, row pitch k_pad); score > 0 only
__global__ void k_binarise(const int32_t* __restrict__ r_user, const int32_t* __restrict__ r_item,
                           const float* __restrict__ r_score, int64_t nnz, int32_t n_user_ids, int32_t n_items,
                           int32_t k_pad, uint8_t* __restrict__ Bt, int* __restrict__ flags) {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= nnz || !(r_score[e] > 0.0f)) return;
    const int32_t u = r_user[e], i = r_item[e];
    if (u < 0 || u >= n_user_ids || i < 0 || i >= n_items) { atomicOr(&flags[DF_BAD_ITEM], 1); return; }
    Bt[(size_t)i * k_pad + u] = 1;
}

// rows [row0, row0+gridDim.x) of the count matrix -> score rows for k_topn: NaN for the item itself
// (excludeSelfSimilarity) and for zero counts (RowSimilarityJob emits no zero similarities)
__global__ void __launch_bounds__(SCORE_THREADS)
k_cooc_scores(const int32_t* __restrict__ C, int32_t n_items, int32_t ldc, int32_t ld, int32_t row0, int32_t self0,
              double* __restrict__ scores, unsigned long long* __restrict__ ustat) {
    const int32_t row = row0 + blockIdx.x;               // row of C
    const int32_t self = self0 + row;                    // the column excluded as "self"
    const int32_t i = blockIdx.y * SCORE_TILE + 2 * threadIdx.x;
    const double NANV = __longlong_as_double(0x7ff8000000000000ll);
    double s0 = NANV, s1 = NANV;
    if (i < n_items && i != self) { const int32_t c = C[(size_t)row * ldc + i]; if (c > 0) s0 = (double)c; }
    if (i + 1 < n_items && i + 1 != self) { const int32_t c = C[(size_t)row * ldc + i + 1]; if (c > 0) s1 = (double)c; }
    double2 out; out.x = s0; out.y = s1;
    *reinterpret_cast<double2*>(scores + (size_t)blockIdx.x * ld + i) = out;
    const bool v0 = (s0 == s0), v1 = (s1 == s1);
    unsigned long long kmin = ~0ull, kmax = 0ull;
    if (v0) { const unsigned long long k = desc_key(s0); kmin = k; kmax = k; }
    if (v1) { const unsigned long long k = desc_key(s1); kmin = min(kmin, k); kmax = max(kmax, k); }
    int cnt = (int)v0 + (int)v1;
    for (int o = 16; o > 0; o >>= 1) {
        cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
        kmin = min(kmin, __shfl_xor_sync(0xffffffffu, kmin, o));
        kmax = max(kmax, __shfl_xor_sync(0xffffffffu, kmax, o));
    }
    if ((threadIdx.x & 31) == 0 && cnt > 0) {
        unsigned long long* st = ustat + 3 * (size_t)blockIdx.x;
        atomicAdd(st, (unsigned long long)cnt);
        atomicMin(st + 1, kmin);
        atomicMax(st + 2, kmax);
    }
}

__global__ void k_iota(int32_t* __restrict__ p, int32_t n) {
    const int32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = i;
}

__global__ void k_widen_counts(const int32_t* __restrict__ cnt, int32_t n, int64_t* __restrict__ out) {
    const int32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = cnt[i];
}

// ---------------------------------------------------------------------------------------------
// Roofline probe: the memory side of k_score_f32 with the arithmetic removed.  CTA = (user, 512-column tile), 128
// threads, one 16-byte __ldg per thread and row, `rows_per_user` rows per CTA in a per-user pseudo-random order over
// an [n_rows x ld] 4-byte plane (54 MB panel per tile column at ML-20M shape: L2 resident, like the real kernel).
// bytes / time of this kernel = what the L2 -> SM path delivers for this access pattern on this box.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(SCORE_THREADS)
k_probe_plane_read(const uint32_t* __restrict__ plane, int32_t n_rows, int32_t ld, int32_t rows_per_user,
                   uint32_t* __restrict__ sink) {
    const uint32_t* __restrict__ base = plane + (size_t)blockIdx.y * SCOREH_TILE + 4 * threadIdx.x;
    uint32_t state = 0x9e3779b9u * (blockIdx.x + 1);
    uint4 acc = make_uint4(0, 0, 0, 0);
    for (int32_t k = 0; k < rows_per_user; k += 8) {
        uint4 h[8];
#pragma unroll
        for (int q = 0; q < 8; q++) {
            state = state * 1664525u + 1013904223u;
            const uint32_t row = (uint32_t)(((uint64_t)state * (uint32_t)n_rows) >> 32);
            h[q] = __ldg(reinterpret_cast<const uint4*>(base + (size_t)row * ld));
        }
#pragma unroll
        for (int q = 0; q < 8; q++) { acc.x ^= h[q].x; acc.y ^= h[q].y; acc.z ^= h[q].z; acc.w ^= h[q].w; }
    }
    if ((acc.x ^ acc.y ^ acc.z ^ acc.w) == 0x12345678u) sink[0] = acc.x;      // keeps the loads alive
}

// packed output (+ per-row user id / cluster, so that a host can rebuild the triples from the 12-byte (item, score) stream)
__global__ void k_pack(const int32_t* __restrict__ out_item, const double* __restrict__ out_score,
                       const int32_t* __restrict__ out_count, const int64_t* __restrict__ out_off,
                       int32_t out_stride, int32_t rank_begin, int32_t n_rows,
                       const int32_t* __restrict__ rank_userid, const int32_t* __restrict__ rank_cluster,
                       int32_t* __restrict__ p_user, int32_t* __restrict__ p_item, double* __restrict__ p_s64,
                       float* __restrict__ p_s32, int32_t* __restrict__ p_cluster,
                       int32_t* __restrict__ row_user, int32_t* __restrict__ row_cluster) {
    const int32_t r = blockIdx.x;
    if (r >= n_rows) return;
    const int32_t n = out_count[r];
    const int64_t off = out_off[r];
    const int32_t rank = rank_begin + r;
    const int32_t uid = rank_userid[rank], c = rank_cluster[rank];
    if (threadIdx.x == 0) { row_user[r] = uid; row_cluster[r] = c; }
    FY_CHECK(n >= 0 && n <= out_stride && off >= 0 && off <= (int64_t)r * out_stride);
    for (int32_t t = threadIdx.x; t < n; t += blockDim.x) {
        const double s = out_score[(size_t)r * out_stride + t];
        p_user[off + t] = uid;
        p_item[off + t] = out_item[(size_t)r * out_stride + t];
        p_s64[off + t] = s;
        p_s32[off + t] = (float)s;                        // RM2HDFSReducer.java:48
        p_cluster[off + t] = c;
    }
}

}  // namespace fy
