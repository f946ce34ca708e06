// rm2_engine.cu -- host side of libfilmyou_rm2.so: context, C ABI (include/filmyou_rm2.h) and the
// stream-ordered pipeline that replaces jobs RM2-1..3 of M/rm/RM2Job.java:76-100.
//
// Pipeline of fy_rm2_run (index phase on the context's stream, then three streams; host sync points: A for input
// errors and sizes, B / B2 for the per-cluster plan, C1 / C2 at the end):
//   index : ratings -> (user rank, item) sort -> CSR; (item, user rank) sort -> CSC; user sums,
//           truncated total, p(i|C); per-cluster local item numbering; d, alpha, c(u,j); with shard_count > 1 and exact
//           (dyadic) scores each rank indexes only the ratings of the clusters it touches
//   per cluster of this shard:  stream G: k_build_H2 (cluster c+1)  |  stream S: k_score_f32 (cluster c)  |
//                               stream T: k_topn + k_refine_score + k_refine_sort (cluster c-1); H and scores double buffered
//   exchange (communicator attached): grouped in-place ncclBroadcast of every rank's [rows x N] block
//   pack  : dense [user x N] results -> packed triples + one (user, cluster, count) record per row
// cub::DeviceRadixSort / DeviceScan / DeviceSegmentedSort are used for the plumbing sorts and scans only; every
// kernel on the scoring path is in rm2_kernels.cuh.
#include "../../include/filmyou_rm2.h"
#include "rm2_kernels.cuh"

#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>
#include <cub/device/device_segmented_sort.cuh>

#include <dlfcn.h>
#include <nccl.h>          // types and prototypes only: libnccl.so.2 is resolved with dlopen at run time

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <numeric>
#include <thread>
#include <vector>

namespace {

struct CudaFail { cudaError_t err; const char* what; int line; };
#define CK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) throw CudaFail{e_, #call, __LINE__}; } while (0)
struct StatusFail { int code; };

template <class T>
struct DBuf {                      // grow-only device buffer
    T* p = nullptr;
    size_t cap = 0;
    void need(size_t n) {
        if (n <= cap) return;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        cudaError_t e = cudaMalloc((void**)&p, std::max<size_t>(n, 1) * sizeof(T));
        if (e != cudaSuccess) { p = nullptr; throw CudaFail{e, "cudaMalloc", __LINE__}; }
        cap = n;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
    ~DBuf() { release(); }
};

inline int bits_for(uint64_t n_values) {   // bits needed to represent 0..n_values-1
    int b = 1;
    while (b < 63 && (1ull << b) < n_values) b++;
    return b;
}
inline int cdiv(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

}  // namespace

struct fy_rm2_ctx {
    fy_rm2_params prm{};
    char err[512] = {0};
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    int64_t launches = 0;

    // ---- inputs ----
    int64_t nnz = 0;
    int32_t max_item = -1;
    DBuf<int32_t> in_user, in_item;
    DBuf<float> in_score;
    bool have_ratings = false, have_clustering = false, have_results = false;

    // clustering (host)
    int32_t n_users = 0, n_clusters = 0;
    std::vector<int32_t> h_rank_userid, h_rank_cluster, h_cstart, h_input_rank;  // input order -> rank
    DBuf<int32_t> uid_sorted, uid_rank, uid_table, rank_userid, rank_cluster, cstart;
    int32_t uid_table_n = 0, uid_table_min = 0;      // direct user id -> rank table (dense ids), else binary search

    // fine-seam overrides (rank order / item id order), empty when unused
    DBuf<double> ext_usum, ext_iprob;
    bool use_ext = false;
    int32_t split = 0, n_splits = 1;

    // ---- index ----
    DBuf<uint64_t> keys_a, keys_b, keys_c;
    DBuf<float> s_score;
    DBuf<int32_t> src_a, csc_src, rowptr;
    DBuf<unsigned char> cub_tmp;
    DBuf<double> usum, isum, iprob, bvec, total, work, work_scan, tsum;
    DBuf<int32_t> n_u, shard_dev;            // shard_dev: bounds[world + 1], rl, rh (k_shard_bounds)
    std::vector<int32_t> h_bounds;           // shard boundaries of every rank (user ranks)
    DBuf<double> run_terms;
    DBuf<float> c_score;
    DBuf<unsigned long long> ustat[2];
    bool exact_scores = false;
    DBuf<unsigned long long> counters;   // [0] n_valid, [1] truncated counter, [2] bits of the smallest positive b_i
    DBuf<int> flags, imax;
    DBuf<int32_t> ifirst, ilast, tstart, tend, tloc, icount, item_off;
    DBuf<int32_t> c_item, c_start, c_len, csr_loc, csc_lu;
    DBuf<double> c_b, c_alpha, csr_delta, csc_delta, csr_c;
    int32_t m = 0;
    std::vector<int32_t> h_icount, h_item_off;
    double h_total = 0.0;

    // ---- per cluster ----
    DBuf<double> H[3], scores[2];           // H: two buffers (FY_H_BUFS=3: three, an experiment)
    DBuf<uint32_t> Hh[3];
    DBuf<int32_t> cand[2], cand_cnt[2];
    DBuf<double> cand_score[2];
    DBuf<int> overflow;
    DBuf<uint64_t> perm_keys[2];
    DBuf<int32_t> perm_vals, perm, row_perm, row_perm_vals;
    DBuf<uint64_t> row_keys[2];
    DBuf<uint64_t> sort_keys[2];
    DBuf<int32_t> sort_idx[2], seg_off;
    DBuf<unsigned char> sort_tmp;
    DBuf<unsigned long long> cbound;
    DBuf<int32_t> chunk_ptr2[3];
    cudaStream_t stream_g = nullptr, stream_t = nullptr;   // H build / top-N run beside the score stream
    cudaEvent_t sync_ev[12] = {nullptr};

    // ---- results ----
    int32_t shard_begin = 0, shard_end = 0, out_stride = 0;
    DBuf<int32_t> out_item, out_count;
    DBuf<double> out_score;
    DBuf<int64_t> out_off, out_cnt64;
    int64_t n_results = 0, users_scored = 0;
    int32_t n_result_rows = 0;
    DBuf<int32_t> row_user, row_cluster;
    void* nccl_comm = nullptr;                      // ncclComm_t, attached by fy_rm2_comm_init
    std::vector<fy_rm2_ctx*> kids;                  // n_gpus > 1: one child context per device
    std::vector<int32_t> owner_rank;                // neighbour-list mode (child context): the one scored rank of each virtual cluster
    fy_rm2_ctx* nbr_child = nullptr;                // lazily created by fy_rm2_run_neighbours
    fy_rm2_ctx* view = nullptr;                     // results of the last call live in this (neighbour-list) child context
    std::vector<int64_t> kid_off;
    DBuf<int32_t> p_user, p_item, p_cluster;
    DBuf<double> p_s64;
    DBuf<float> p_s32;

    // ---- co-occurrence (config 3) ----
    DBuf<int32_t> cooc_counts, cooc_iota, cooc_zero, cooc_out_item, cooc_out_cnt;
    DBuf<uint8_t> cooc_bt;
    DBuf<double> cooc_out_score;
    int32_t cooc_items = 0, cooc_ldc = 0;

    fy_rm2_profile prof{};
    std::vector<cudaEvent_t> events;

    int fail(int code, const char* fmt, ...) {
        va_list ap; va_start(ap, fmt); vsnprintf(err, sizeof(err), fmt, ap); va_end(ap);
        return code;
    }
    cudaEvent_t ev(size_t i) {
        while (events.size() <= i) { cudaEvent_t e; CK(cudaEventCreate(&e)); events.push_back(e); }
        return events[i];
    }
};

#define LAUNCH(ctx, kernel, grid, block, smem, ...)                                   \
    do {                                                                              \
        kernel<<<(grid), (block), (smem), (ctx)->stream>>>(__VA_ARGS__);              \
        (ctx)->launches++;                                                            \
        CK(cudaGetLastError());                                                       \
    } while (0)

#define LAUNCH_ON(ctx, strm, kernel, grid, block, smem, ...)                          \
    do {                                                                              \
        kernel<<<(grid), (block), (smem), (strm)>>>(__VA_ARGS__);                     \
        (ctx)->launches++;                                                            \
        CK(cudaGetLastError());                                                       \
    } while (0)

// runs f, turning C++/CUDA failures into fy_status codes: nothing crosses the C boundary
template <class F>
static int guarded(fy_rm2_ctx* ctx, F&& f) {
    try {
        return f();
    } catch (const CudaFail& c) {
        const int code = (c.err == cudaErrorMemoryAllocation) ? FY_E_NOMEM : FY_E_CUDA;
        return ctx->fail(code, "CUDA error %d (%s) at %s, rm2_engine.cu:%d", (int)c.err, cudaGetErrorString(c.err), c.what, c.line);
    } catch (const StatusFail& s) {
        return s.code;
    } catch (const std::bad_alloc&) {
        return ctx->fail(FY_E_NOMEM, "host allocation failed");
    } catch (...) {
        return ctx->fail(FY_E_CUDA, "unexpected exception");
    }
}

// n_gpus > 1: run f on every child context, one host thread per device (each call has its own host sync points)
template <class F>
static int fan_out(fy_rm2_ctx* ctx, F&& f) {
    const size_t n = ctx->kids.size();
    std::vector<int> rc(n, FY_OK);
    std::vector<std::thread> th;
    bool started_all = true;
    try {
        th.reserve(n);
        for (size_t i = 0; i < n; i++) th.emplace_back([&, i]() { rc[i] = f(ctx->kids[i], i); });
    } catch (...) {
        started_all = false;           // join what did start before reporting: a joinable std::thread must not be destroyed
    }
    for (std::thread& t : th) t.join();
    if (!started_all) return ctx->fail(FY_E_NOMEM, "could not start a host thread per device");
    for (size_t i = 0; i < n; i++)
        if (rc[i] != FY_OK) return ctx->fail(rc[i], "device %d: %s", ctx->kids[i]->prm.device, ctx->kids[i]->err);
    return FY_OK;
}

extern "C" int fy_rm2_abi_version(void) { return FY_RM2_ABI_VERSION; }

#ifdef FY_BOUNDS_CHECK
// checked build only (tools/build_checked.py): device-side range violations counted since the library was loaded
extern "C" long long fy_rm2_debug_violations(void) {
    unsigned long long v = 0;
    cudaDeviceSynchronize();
    if (cudaMemcpyFromSymbol(&v, fy::g_bounds_violations, sizeof(v)) != cudaSuccess) return -1;
    return (long long)v;
}
#endif

extern "C" void fy_rm2_default_params(fy_rm2_params* p) {
    if (!p) return;
    std::memset(p, 0, sizeof(*p));
    p->lambda = 0.1;                 // RMRecommenderDriver.java:114
    p->number_of_items = 0;          // required (:91)
    p->top_n = 1000;                 // :95
    p->filter_users = 0;             // :119
    p->device = 0;
    p->shard_rank = 0;
    p->shard_count = 1;
    p->tie_break = 0;
    p->score_mode = 0;
    p->n_gpus = 0;
    p->reserved = 0;
}

extern "C" int fy_rm2_create(fy_rm2_ctx** out, const fy_rm2_params* p) {
    if (!out || !p) return FY_E_ARG;
    *out = nullptr;
    if (p->top_n < 0 || p->number_of_items <= 0 || p->tie_break != 0 || p->score_mode < 0 || p->score_mode > 1 || !(p->lambda >= 0.0 && p->lambda <= 1.0))
        return FY_E_ARG;
    if (p->shard_count < 0 || (p->shard_count > 1 && (p->shard_rank < 0 || p->shard_rank >= p->shard_count)))
        return FY_E_ARG;
    int n_dev = 0;
    if (cudaGetDeviceCount(&n_dev) != cudaSuccess || n_dev <= 0 || p->device < 0 || p->device >= n_dev)
        return FY_E_CUDA;            // no CPU fallback: fail loudly
    if (p->n_gpus < 0 || p->reserved != 0) return FY_E_ARG;
    if (p->n_gpus > 1 && (p->shard_count > 1 || p->device + p->n_gpus > n_dev)) return FY_E_ARG;
    fy_rm2_ctx* ctx = new (std::nothrow) fy_rm2_ctx();
    if (!ctx) return FY_E_NOMEM;
    ctx->prm = *p;
    if (ctx->prm.shard_count <= 0) { ctx->prm.shard_count = 1; ctx->prm.shard_rank = 0; }
    if (p->n_gpus > 1) {                 // one child context per device; the parent only fans calls out
        for (int i = 0; i < p->n_gpus; i++) {
            fy_rm2_params kp = *p;
            kp.device = p->device + i; kp.shard_rank = i; kp.shard_count = p->n_gpus; kp.n_gpus = 0;
            fy_rm2_ctx* kid = nullptr;
            const int rc = fy_rm2_create(&kid, &kp);
            if (rc != FY_OK) { fy_rm2_destroy(ctx); return rc; }
            ctx->kids.push_back(kid);
        }
        *out = ctx;
        return FY_OK;
    }
    int rc = guarded(ctx, [&]() {
        CK(cudaSetDevice(p->device));
        cudaDeviceProp prop;
        CK(cudaGetDeviceProperties(&prop, p->device));
        if (prop.major < 10) return ctx->fail(FY_E_CUDA, "device %d is sm_%d%d; this library is built for sm_100a only", p->device, prop.major, prop.minor);
        CK(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
        ctx->own_stream = true;
        return (int)FY_OK;
    });
    if (rc != FY_OK) { fy_rm2_destroy(ctx); return rc; }
    *out = ctx;
    return FY_OK;
}

extern "C" void fy_rm2_destroy(fy_rm2_ctx* ctx) {
    if (!ctx) return;
    for (fy_rm2_ctx* k : ctx->kids) fy_rm2_destroy(k);
    ctx->kids.clear();
    if (ctx->nbr_child) { fy_rm2_destroy(ctx->nbr_child); ctx->nbr_child = nullptr; }
    fy_rm2_comm_destroy(ctx);
    cudaSetDevice(ctx->prm.device);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    for (cudaEvent_t e : ctx->events) cudaEventDestroy(e);
    for (cudaEvent_t e : ctx->sync_ev) if (e) cudaEventDestroy(e);
    if (ctx->stream_g) { cudaStreamSynchronize(ctx->stream_g); cudaStreamDestroy(ctx->stream_g); }
    if (ctx->stream_t) { cudaStreamSynchronize(ctx->stream_t); cudaStreamDestroy(ctx->stream_t); }
    if (ctx->own_stream && ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

extern "C" const char* fy_rm2_last_error(const fy_rm2_ctx* ctx) { return ctx ? ctx->err : "null context"; }

extern "C" int fy_rm2_set_stream(fy_rm2_ctx* ctx, void* cuda_stream) {
    if (!ctx) return FY_E_ARG;
    if (!ctx->kids.empty()) return ctx->fail(FY_E_UNSUPPORTED, "fy_rm2_set_stream on an n_gpus > 1 context (each device runs on its own stream)");
    return guarded(ctx, [&]() {
        CK(cudaSetDevice(ctx->prm.device));
        if (ctx->stream) CK(cudaStreamSynchronize(ctx->stream));
        if (ctx->own_stream && ctx->stream) CK(cudaStreamDestroy(ctx->stream));
        ctx->stream = (cudaStream_t)cuda_stream;
        ctx->own_stream = false;
        return (int)FY_OK;
    });
}

static int check_flags(fy_rm2_ctx* ctx, const int* f) {
    if (f[fy::DF_BAD_ITEM]) return ctx->fail(FY_E_ARG, "negative or out-of-range item id in the ratings");
    if (f[fy::DF_UNKNOWN_USER]) return ctx->fail(FY_E_UNKNOWN_USER, "a positive rating belongs to a user absent from `clustering`");
    if (f[fy::DF_DUPLICATE]) return ctx->fail(FY_E_DUPLICATE_RATING, "the same (user,item) pair is rated twice");
    if (f[fy::DF_USER_WITHOUT_RATING]) return ctx->fail(FY_E_USER_WITHOUT_RATING, "`clustering` lists a user without any positive rating (AbstractRM2Reducer.java:153-160 would mis-parse the group)");
    return FY_OK;
}

static int upload_ratings(fy_rm2_ctx* ctx, const int32_t* user, const int32_t* item, const float* score, int64_t nnz) {
    if (nnz > 0x7fffffffll - 1024) return ctx->fail(FY_E_UNSUPPORTED, "more than 2^31 ratings");
    CK(cudaSetDevice(ctx->prm.device));
    ctx->have_ratings = false;
    ctx->have_results = false;
    ctx->in_user.need((size_t)nnz); ctx->in_item.need((size_t)nnz); ctx->in_score.need((size_t)nnz);
    ctx->flags.need(fy::DF_COUNT); ctx->imax.need(1);
    cudaStream_t st = ctx->stream;
    CK(cudaMemcpyAsync(ctx->in_user.p, user, (size_t)nnz * 4, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(ctx->in_item.p, item, (size_t)nnz * 4, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(ctx->in_score.p, score, (size_t)nnz * 4, cudaMemcpyHostToDevice, st));
    CK(cudaMemsetAsync(ctx->flags.p, 0, sizeof(int) * fy::DF_COUNT, st));
    CK(cudaMemsetAsync(ctx->imax.p, 0xff, sizeof(int), st));
    if (nnz > 0) LAUNCH(ctx, fy::k_scan_ratings, cdiv(nnz, 256), 256, 0, ctx->in_item.p, ctx->in_score.p, nnz, ctx->imax.p, ctx->flags.p);
    int h_flags[fy::DF_COUNT]; int h_max = -1;
    CK(cudaMemcpyAsync(h_flags, ctx->flags.p, sizeof(h_flags), cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(&h_max, ctx->imax.p, sizeof(int), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    int rc = check_flags(ctx, h_flags);
    if (rc != FY_OK) return rc;
    if (h_max < 0) return ctx->fail(FY_E_ARG, "no rating with score > 0");
    ctx->nnz = nnz;
    ctx->exact_scores = (h_flags[fy::DF_INEXACT_SCORES] == 0);
    ctx->max_item = h_max;
    ctx->have_ratings = true;
    return FY_OK;
}

extern "C" int fy_rm2_set_ratings(fy_rm2_ctx* ctx, const int32_t* user, const int32_t* item, const float* score, int64_t nnz) {
    if (!ctx) return FY_E_ARG;
    if (!user || !item || !score || nnz <= 0) return ctx->fail(FY_E_ARG, "fy_rm2_set_ratings: null pointer or nnz <= 0");
    if (!ctx->kids.empty()) {
        ctx->have_results = false;
        return fan_out(ctx, [&](fy_rm2_ctx* k, size_t) { return fy_rm2_set_ratings(k, user, item, score, nnz); });
    }
    ctx->view = nullptr;
    return guarded(ctx, [&]() { ctx->use_ext = false; return upload_ratings(ctx, user, item, score, nnz); });
}

static int upload_clustering(fy_rm2_ctx* ctx, const int32_t* user, const int32_t* cluster, int64_t n_users,
                             const int32_t* cluster_size, int32_t n_clusters, bool check_sizes) {
    if (n_users > 0x7ffffff0ll) return ctx->fail(FY_E_UNSUPPORTED, "too many users");
    ctx->have_clustering = false;
    ctx->have_results = false;
    const int32_t U = (int32_t)n_users;
    std::vector<int32_t> cnt((size_t)n_clusters, 0);
    for (int32_t k = 0; k < U; k++) {
        if (cluster[k] < 0 || cluster[k] >= n_clusters) return ctx->fail(FY_E_ARG, "cluster id %d of user %d outside [0,%d)", cluster[k], user[k], n_clusters);
        cnt[cluster[k]]++;
    }
    if (check_sizes)
        for (int32_t c = 0; c < n_clusters; c++)
            if (cnt[c] != cluster_size[c])
                return ctx->fail(FY_E_CLUSTER_SIZE, "clusteringCount[%d] = %d but %d users map to it", c, cluster_size[c], cnt[c]);
    // users[] of the reducer, canonicalised: (cluster, user id) ascending.  A stable counting sort by cluster
    // (O(U)) leaves each cluster's users in input order; only segments whose ids are not already ascending
    // (the `clustering` file is normally written by ascending user id) pay for a comparison sort.
    ctx->h_cstart.assign((size_t)n_clusters + 1, 0);
    for (int32_t c = 0; c < n_clusters; c++) ctx->h_cstart[c + 1] = ctx->h_cstart[c] + cnt[c];
    std::vector<int32_t> order((size_t)U);
    {
        std::vector<int32_t> fill(ctx->h_cstart.begin(), ctx->h_cstart.end() - 1);
        for (int32_t k = 0; k < U; k++) order[fill[cluster[k]]++] = k;
        for (int32_t c = 0; c < n_clusters; c++) {
            int32_t* b = order.data() + ctx->h_cstart[c]; int32_t* e = order.data() + ctx->h_cstart[c + 1];
            bool sorted = true;
            for (int32_t* q = b; q + 1 < e; q++) if (user[q[0]] >= user[q[1]]) { sorted = false; break; }
            if (!sorted) std::sort(b, e, [&](int32_t a, int32_t bb) { return user[a] < user[bb]; });
        }
    }
    ctx->h_rank_userid.resize(U); ctx->h_rank_cluster.resize(U); ctx->h_input_rank.resize(U);
    int32_t uid_min = 0x7fffffff, uid_max = -0x7fffffff - 1;
    for (int32_t r = 0; r < U; r++) {
        const int32_t id = user[order[r]];
        ctx->h_rank_userid[r] = id;
        ctx->h_rank_cluster[r] = cluster[order[r]];
        ctx->h_input_rank[order[r]] = r;
        uid_min = std::min(uid_min, id); uid_max = std::max(uid_max, id);
    }
    CK(cudaSetDevice(ctx->prm.device));
    cudaStream_t st = ctx->stream;
    // user id -> rank: a direct table when the ids are dense enough (one load per rating in k_make_keys), else
    // the id-sorted arrays that k_make_keys binary-searches
    std::vector<int32_t> ids, ranks, table;
    const int64_t span = (int64_t)uid_max - (int64_t)uid_min + 1;
    ctx->uid_table_n = 0; ctx->uid_table_min = 0;
    if (uid_min >= 0 && span <= 8ll * U + (1ll << 20)) {
        table.assign((size_t)span, -1);
        for (int32_t r = 0; r < U; r++) {
            int32_t& slot = table[(size_t)(ctx->h_rank_userid[r] - uid_min)];
            if (slot >= 0) return ctx->fail(FY_E_ARG, "user %d appears twice in `clustering`", ctx->h_rank_userid[r]);
            slot = r;
        }
        ctx->uid_table.need((size_t)span);
        CK(cudaMemcpyAsync(ctx->uid_table.p, table.data(), (size_t)span * 4, cudaMemcpyHostToDevice, st));
        ctx->uid_table_n = (int32_t)span; ctx->uid_table_min = uid_min;
    } else {
        std::vector<int32_t> by_id((size_t)U);
        std::iota(by_id.begin(), by_id.end(), 0);
        std::sort(by_id.begin(), by_id.end(), [&](int32_t a, int32_t b) { return ctx->h_rank_userid[a] < ctx->h_rank_userid[b]; });
        ids.resize((size_t)U); ranks.resize((size_t)U);
        for (int32_t k = 0; k < U; k++) { ids[k] = ctx->h_rank_userid[by_id[k]]; ranks[k] = by_id[k]; }
        for (int32_t k = 1; k < U; k++)
            if (ids[k] == ids[k - 1]) return ctx->fail(FY_E_ARG, "user %d appears twice in `clustering`", ids[k]);
        ctx->uid_sorted.need(U); ctx->uid_rank.need(U);
        CK(cudaMemcpyAsync(ctx->uid_sorted.p, ids.data(), (size_t)U * 4, cudaMemcpyHostToDevice, st));
        CK(cudaMemcpyAsync(ctx->uid_rank.p, ranks.data(), (size_t)U * 4, cudaMemcpyHostToDevice, st));
    }
    ctx->rank_userid.need(U); ctx->rank_cluster.need(U);
    ctx->cstart.need((size_t)n_clusters + 1);
    CK(cudaMemcpyAsync(ctx->rank_userid.p, ctx->h_rank_userid.data(), (size_t)U * 4, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(ctx->rank_cluster.p, ctx->h_rank_cluster.data(), (size_t)U * 4, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(ctx->cstart.p, ctx->h_cstart.data(), ((size_t)n_clusters + 1) * 4, cudaMemcpyHostToDevice, st));
    CK(cudaStreamSynchronize(st));   // the staging vectors die at return
    ctx->n_users = U;
    ctx->n_clusters = n_clusters;
    ctx->have_clustering = true;
    return FY_OK;
}

extern "C" int fy_rm2_set_clustering(fy_rm2_ctx* ctx, const int32_t* user, const int32_t* cluster, int64_t n_users,
                                     const int32_t* cluster_size, int32_t n_clusters) {
    if (!ctx) return FY_E_ARG;
    if (!user || !cluster || !cluster_size || n_users <= 0 || n_clusters <= 0)
        return ctx->fail(FY_E_ARG, "fy_rm2_set_clustering: null pointer or empty input");
    ctx->view = nullptr;
    if (!ctx->kids.empty()) {
        ctx->have_results = false;
        ctx->n_users = (int32_t)n_users;
        return fan_out(ctx, [&](fy_rm2_ctx* k, size_t) { return fy_rm2_set_clustering(k, user, cluster, n_users, cluster_size, n_clusters); });
    }
    return guarded(ctx, [&]() { ctx->use_ext = false; ctx->n_splits = 1; ctx->split = 0;
                                return upload_clustering(ctx, user, cluster, n_users, cluster_size, n_clusters, true); });
}

// ---------------------------------------------------------------------------------------------
// NCCL, resolved at run time (the process may already hold torch's bundled libnccl.so.2: same soname, same handle)
// ---------------------------------------------------------------------------------------------
namespace {
struct NcclApi {
    void* lib = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*Broadcast)(const void*, void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    bool ok = false;
};
NcclApi& nccl_api() {
    static NcclApi api = [] {
        NcclApi a;
        a.lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!a.lib) a.lib = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
        if (!a.lib) return a;
        a.GetUniqueId = (decltype(a.GetUniqueId))dlsym(a.lib, "ncclGetUniqueId");
        a.CommInitRank = (decltype(a.CommInitRank))dlsym(a.lib, "ncclCommInitRank");
        a.CommDestroy = (decltype(a.CommDestroy))dlsym(a.lib, "ncclCommDestroy");
        a.Broadcast = (decltype(a.Broadcast))dlsym(a.lib, "ncclBroadcast");
        a.GroupStart = (decltype(a.GroupStart))dlsym(a.lib, "ncclGroupStart");
        a.GroupEnd = (decltype(a.GroupEnd))dlsym(a.lib, "ncclGroupEnd");
        a.GetErrorString = (decltype(a.GetErrorString))dlsym(a.lib, "ncclGetErrorString");
        a.ok = a.GetUniqueId && a.CommInitRank && a.CommDestroy && a.Broadcast && a.GroupStart && a.GroupEnd;
        return a;
    }();
    return api;
}
}  // namespace

// every rank's [rows x stride] block of (item, score) and its row counts, broadcast in place from its owner
static int nccl_exchange(fy_rm2_ctx* ctx, int32_t stride, cudaStream_t st) {
    NcclApi& n = nccl_api();
    if (!n.ok) return ctx->fail(FY_E_UNSUPPORTED, "libnccl.so.2 not available");
    ncclComm_t comm = (ncclComm_t)ctx->nccl_comm;
    const int world = ctx->prm.shard_count;
    auto chk = [&](ncclResult_t r, const char* what) {
        if (r != ncclSuccess) { ctx->fail(FY_E_CUDA, "NCCL error %d (%s) in %s", (int)r, n.GetErrorString ? n.GetErrorString(r) : "?", what); throw StatusFail{FY_E_CUDA}; }
    };
    chk(n.GroupStart(), "ncclGroupStart");
    for (int r = 0; r < world; r++) {
        const int32_t r0 = ctx->h_bounds[r], rows = ctx->h_bounds[r + 1] - r0;
        if (rows <= 0) continue;
        int32_t* pi = ctx->out_item.p + (size_t)r0 * stride;
        double* ps = ctx->out_score.p + (size_t)r0 * stride;
        int32_t* pc = ctx->out_count.p + r0;
        chk(n.Broadcast(pi, pi, (size_t)rows * stride, ncclInt32, r, comm, st), "ncclBroadcast(items)");
        chk(n.Broadcast(ps, ps, (size_t)rows * stride, ncclFloat64, r, comm, st), "ncclBroadcast(scores)");
        chk(n.Broadcast(pc, pc, (size_t)rows, ncclInt32, r, comm, st), "ncclBroadcast(counts)");
    }
    chk(n.GroupEnd(), "ncclGroupEnd");
    return FY_OK;
}

extern "C" int fy_rm2_nccl_unique_id(void* id_out) {
    if (!id_out) return FY_E_ARG;
    NcclApi& n = nccl_api();
    if (!n.ok) return FY_E_UNSUPPORTED;
    static_assert(sizeof(ncclUniqueId) == FY_NCCL_UNIQUE_ID_BYTES, "ncclUniqueId size");
    ncclUniqueId id;
    if (n.GetUniqueId(&id) != ncclSuccess) return FY_E_CUDA;
    std::memcpy(id_out, &id, sizeof(id));
    return FY_OK;
}

extern "C" int fy_rm2_comm_destroy(fy_rm2_ctx* ctx) {
    if (!ctx) return FY_E_ARG;
    if (ctx->nccl_comm) {
        cudaSetDevice(ctx->prm.device);
        if (ctx->stream) cudaStreamSynchronize(ctx->stream);
        nccl_api().CommDestroy((ncclComm_t)ctx->nccl_comm);
        ctx->nccl_comm = nullptr;
    }
    return FY_OK;
}

extern "C" int fy_rm2_comm_init(fy_rm2_ctx* ctx, const void* id, int32_t world, int32_t rank) {
    if (!ctx) return FY_E_ARG;
    if (!id || world < 1 || rank < 0 || rank >= world) return ctx->fail(FY_E_ARG, "fy_rm2_comm_init: bad argument");
    if (!ctx->kids.empty()) return ctx->fail(FY_E_UNSUPPORTED, "fy_rm2_comm_init on an n_gpus > 1 context");
    if (ctx->prm.shard_count != world || ctx->prm.shard_rank != rank)
        return ctx->fail(FY_E_ARG, "communicator (rank %d of %d) does not match the context's shard (%d of %d)", rank, world, ctx->prm.shard_rank, ctx->prm.shard_count);
    NcclApi& n = nccl_api();
    if (!n.ok) return ctx->fail(FY_E_UNSUPPORTED, "libnccl.so.2 not available (dlopen failed)");
    fy_rm2_comm_destroy(ctx);
    return guarded(ctx, [&]() {
        CK(cudaSetDevice(ctx->prm.device));
        ncclUniqueId uid;
        std::memcpy(&uid, id, sizeof(uid));
        ncclComm_t comm = nullptr;
        const ncclResult_t r = n.CommInitRank(&comm, world, uid, rank);
        if (r != ncclSuccess) return ctx->fail(FY_E_CUDA, "ncclCommInitRank failed: %d (%s)", (int)r, n.GetErrorString ? n.GetErrorString(r) : "?");
        ctx->nccl_comm = (void*)comm;
        return (int)FY_OK;
    });
}

extern "C" int fy_rm2_shard_bounds(const fy_rm2_ctx* ctx, int32_t* bounds) {
    if (!ctx || !bounds) return FY_E_ARG;
    if (!ctx->have_results || ctx->h_bounds.empty()) return FY_E_STATE;
    for (size_t r = 0; r < ctx->h_bounds.size(); r++) bounds[r] = ctx->h_bounds[r];
    return FY_OK;
}

// ---------------------------------------------------------------------------------------------
// the pipeline
// ---------------------------------------------------------------------------------------------
template <int L>
static void launch_score(fy_rm2_ctx* ctx, cudaStream_t strm, dim3 grid, const double* H, int32_t I_c, int32_t ld,
                         int32_t rank_begin, int32_t slot0, double log_items, double log_K, double* scores,
                         unsigned long long* ustat) {
    LAUNCH_ON(ctx, strm, fy::k_score<L>, grid, fy::SCORE_THREADS, 0, H, I_c, ld, rank_begin, slot0, ctx->rowptr.p,
              ctx->csr_loc.p, ctx->csr_c.p, ctx->c_b.p, log_items, log_K, scores, ustat);
}

static int run_pipeline_impl(fy_rm2_ctx* ctx, bool force_exact) {
    using namespace fy;
    if (!ctx->have_ratings || !ctx->have_clustering) return ctx->fail(FY_E_STATE, "fy_rm2_run needs fy_rm2_set_ratings and fy_rm2_set_clustering first");
    CK(cudaSetDevice(ctx->prm.device));
    cudaStream_t st = ctx->stream;
    ctx->have_results = false;
    ctx->launches = 0;
    ctx->prof = fy_rm2_profile{};
    const int64_t nnz = ctx->nnz;
    const int32_t U = ctx->n_users, KC = ctx->n_clusters;
    const int32_t TI = ctx->max_item + 1;                    // direct item tables
    if ((int64_t)KC * TI > (1ll << 28)) return ctx->fail(FY_E_UNSUPPORTED, "n_clusters * (max_item+1) = %lld exceeds the direct-table limit 2^28", (long long)KC * TI);
    const int item_bits = bits_for((uint64_t)TI), rank_bits = bits_for((uint64_t)U);
    const int key_bits = std::min(64, item_bits + rank_bits + 1);
    const double lambda = ctx->prm.lambda;
    size_t evi = 0;
    cudaEvent_t ev_start = ctx->ev(evi++), ev_index = ctx->ev(evi++), ev_end = ctx->ev(evi++);
    CK(cudaEventRecord(ev_start, st));

    // ---------------- index: sort #1 (user rank, item) -> CSR ----------------
    ctx->keys_a.need((size_t)nnz); ctx->keys_b.need((size_t)nnz); ctx->s_score.need((size_t)nnz);
    ctx->counters.need(4); ctx->flags.need(DF_COUNT);
    CK(cudaMemsetAsync(ctx->counters.p, 0, 4 * sizeof(unsigned long long), st));
    CK(cudaMemsetAsync(ctx->counters.p + 2, 0xff, sizeof(unsigned long long), st));
    CK(cudaMemsetAsync(ctx->flags.p, 0, sizeof(int) * DF_COUNT, st));
    const size_t tab = (size_t)KC * TI;
    ctx->rowptr.need((size_t)U + 1);
    ctx->usum.need(U);
    ctx->isum.need(TI); ctx->iprob.need(TI); ctx->bvec.need(TI); ctx->total.need(1);
    ctx->icount.need(KC); ctx->item_off.need((size_t)KC + 1);
    ctx->tstart.need(tab);
    unsigned long long h_counters[4] = {0, 0, 0, 0}; int h_flags[DF_COUNT];
    // Sharded index: with exact (dyadic) scores the global statistics need no sort, so each rank sorts and
    // indexes only the ratings of the clusters it touches (~1/N of them) instead of all of them.
    const int world = ctx->prm.shard_count, me = ctx->prm.shard_rank;
    // share of a cluster's H build in its score work, for the work partition (partition_targets); FY_SHARD_BETA=0 = equal score work
    const char* beta_env = std::getenv("FY_SHARD_BETA");
    const double shard_beta = beta_env ? std::atof(beta_env) : 0.55;
    const bool exact_ok = ctx->exact_scores && U <= (1 << 22) && TI <= (1 << 22);     // <= 2^22 addends per sum (k_scan_ratings)
    const bool sharded_index = world > 1 && exact_ok && !ctx->use_ext;
    int32_t ub = 0, ue = U;
    std::vector<int32_t> h_icount_global((size_t)KC, 0);
    uint64_t* k_sorted = ctx->keys_b.p;       // (rank, item)-sorted CSR keys
    uint64_t* k_scratch = ctx->keys_a.p;
    int32_t m = 0;
    ctx->h_bounds.assign((size_t)world + 1, 0);
    ctx->h_bounds[world] = U;
    unsigned long long h_bmin_bits = ~0ull;
    int32_t rl_h = 0, rh_h = U;               // rank range of the clusters this rank touches (sharded index)
    if (sharded_index) {
        ctx->n_u.need(U); ctx->work.need(U); ctx->work_scan.need(U); ctx->c_score.need((size_t)nnz);
        ctx->shard_dev.need((size_t)world + 3);
        CK(cudaMemsetAsync(ctx->usum.p, 0, (size_t)U * 8, st));
        CK(cudaMemsetAsync(ctx->n_u.p, 0, (size_t)U * 4, st));
        CK(cudaMemsetAsync(ctx->isum.p, 0, (size_t)TI * 8, st));
        CK(cudaMemsetAsync(ctx->tstart.p, 0xff, tab * 4, st));
        // one pass over the replicated ratings: sort keys + user sums, item sums, per-cluster item presence
        LAUNCH(ctx, k_make_keys, cdiv(nnz, 256), 256, 0, ctx->in_user.p, ctx->in_item.p, ctx->in_score.p, nnz,
               ctx->uid_sorted.p, ctx->uid_rank.p, U, ctx->uid_table_n ? ctx->uid_table.p : (const int32_t*)nullptr, ctx->uid_table_n,
               ctx->uid_table_min, item_bits, ctx->max_item, ctx->keys_a.p, ctx->counters.p, ctx->flags.p,
               ctx->rank_cluster.p, TI, ctx->usum.p, ctx->n_u.p, ctx->isum.p, ctx->tstart.p);
        LAUNCH(ctx, k_total_from_usum, cdiv(U, 256), 256, 0, ctx->usum.p, ctx->n_u.p, U, ctx->counters.p + 1, ctx->flags.p);
        LAUNCH(ctx, k_cluster_item_count, KC, 256, 0, ctx->tstart.p, TI, ctx->icount.p);
        LAUNCH(ctx, k_user_work_n, cdiv(U, 256), 256, 0, ctx->n_u.p, ctx->rank_cluster.p, ctx->icount.p, U, ctx->work.p);
        {
            size_t tmp = 0;
            CK(cub::DeviceScan::InclusiveSum(nullptr, tmp, ctx->work.p, ctx->work_scan.p, U, st));
            ctx->cub_tmp.need(tmp);
            CK(cub::DeviceScan::InclusiveSum(ctx->cub_tmp.p, tmp, ctx->work.p, ctx->work_scan.p, U, st));
        }
        // this rank's user range and the rank range of the clusters it touches, computed on the device
        LAUNCH(ctx, k_shard_bounds, 1, std::max(128, ((world + 1 + 31) / 32) * 32), 0, ctx->work_scan.p, U, world, me,
               ctx->rank_cluster.p, ctx->cstart.p, KC, shard_beta, ctx->shard_dev.p);
        LAUNCH(ctx, k_compact_local, cdiv(nnz, 256), 256, 0, ctx->keys_a.p, ctx->in_score.p, nnz, item_bits,
               ctx->shard_dev.p + world + 1, ctx->keys_b.p, ctx->c_score.p, ctx->counters.p + 3);
        LAUNCH(ctx, k_item_prob_isum, cdiv(TI, 128), 128, 0, ctx->isum.p, TI, ctx->counters.p + 1, lambda,
               ctx->iprob.p, ctx->bvec.p, ctx->total.p, ctx->counters.p + 2);
        // sync A (the only one of the index phase): input errors, global item counts, every rank's range, local rating count
        std::vector<int32_t> h_shard((size_t)world + 3);
        CK(cudaMemcpyAsync(h_counters, ctx->counters.p, sizeof(h_counters), cudaMemcpyDeviceToHost, st));
        CK(cudaMemcpyAsync(h_flags, ctx->flags.p, sizeof(h_flags), cudaMemcpyDeviceToHost, st));
        CK(cudaMemcpyAsync(h_icount_global.data(), ctx->icount.p, (size_t)KC * 4, cudaMemcpyDeviceToHost, st));
        CK(cudaMemcpyAsync(h_shard.data(), ctx->shard_dev.p, h_shard.size() * 4, cudaMemcpyDeviceToHost, st));
        CK(cudaMemcpyAsync(&ctx->h_total, ctx->total.p, sizeof(double), cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        { int rc = check_flags(ctx, h_flags); if (rc != FY_OK) return rc; }
        if (h_counters[0] == 0) return ctx->fail(FY_E_ARG, "no positive rating");
        for (int r = 0; r <= world; r++) ctx->h_bounds[r] = h_shard[r];
        ub = h_shard[me]; ue = h_shard[me + 1];
        rl_h = h_shard[world + 1]; rh_h = h_shard[world + 2];
        h_bmin_bits = h_counters[2];
        m = (int32_t)h_counters[3];
        if (m > 0) {
            size_t tmp = 0;
            CK(cub::DeviceRadixSort::SortPairs(nullptr, tmp, ctx->keys_b.p, ctx->keys_a.p, ctx->c_score.p, ctx->s_score.p,
                                               (int64_t)m, 0, key_bits, st));
            ctx->cub_tmp.need(tmp);
            CK(cub::DeviceRadixSort::SortPairs(ctx->cub_tmp.p, tmp, ctx->keys_b.p, ctx->keys_a.p, ctx->c_score.p, ctx->s_score.p,
                                               (int64_t)m, 0, key_bits, st));
        }
        k_sorted = ctx->keys_a.p; k_scratch = ctx->keys_b.p;
    } else {
        LAUNCH(ctx, k_make_keys, cdiv(nnz, 256), 256, 0, ctx->in_user.p, ctx->in_item.p, ctx->in_score.p, nnz,
               ctx->uid_sorted.p, ctx->uid_rank.p, U, ctx->uid_table_n ? ctx->uid_table.p : (const int32_t*)nullptr, ctx->uid_table_n,
               ctx->uid_table_min, item_bits, ctx->max_item, ctx->keys_a.p, ctx->counters.p, ctx->flags.p,
               (const int32_t*)nullptr, TI, (double*)nullptr, (int32_t*)nullptr, (double*)nullptr, (int32_t*)nullptr);
        size_t tmp = 0;
        CK(cub::DeviceRadixSort::SortPairs(nullptr, tmp, ctx->keys_a.p, ctx->keys_b.p, ctx->in_score.p, ctx->s_score.p,
                                           (int64_t)nnz, 0, key_bits, st));
        ctx->cub_tmp.need(tmp);
        CK(cub::DeviceRadixSort::SortPairs(ctx->cub_tmp.p, tmp, ctx->keys_a.p, ctx->keys_b.p, ctx->in_score.p, ctx->s_score.p,
                                           (int64_t)nnz, 0, key_bits, st));
        // sync A: number of positive ratings, input errors
        CK(cudaMemcpyAsync(h_counters, ctx->counters.p, 2 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
        CK(cudaMemcpyAsync(h_flags, ctx->flags.p, sizeof(h_flags), cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        { int rc = check_flags(ctx, h_flags); if (rc != FY_OK) return rc; }
        m = (int32_t)h_counters[0];
        if (m <= 0) return ctx->fail(FY_E_ARG, "no positive rating");
    }
    ctx->m = m;
    const uint64_t* keys = k_sorted;                          // sorted (rank, item)
    const int32_t m1 = std::max(m, 1);

    CK(cudaMemsetAsync(ctx->rowptr.p, 0xff, ((size_t)U + 1) * 4, st));
    if (m > 0) LAUNCH(ctx, k_rows, cdiv(m, 256), 256, 0, keys, m, item_bits, U, ctx->rowptr.p, ctx->flags.p);
    LAUNCH(ctx, k_rows_gap, cdiv(U + 1, 256), 256, 0, keys, m, item_bits, U, ctx->rowptr.p);
    if (!sharded_index)
        LAUNCH(ctx, k_user_sum, cdiv(U, 128), 128, 0, ctx->rowptr.p, ctx->s_score.p, U,
               ctx->use_ext ? ctx->ext_usum.p : (const double*)nullptr, ctx->usum.p, ctx->counters.p + 1, ctx->flags.p);

    // ---------------- sort #2 (item, user rank) -> CSC ----------------
    ctx->src_a.need(m1); ctx->csc_src.need(m1); ctx->keys_c.need(m1);
    if (m > 0) {
        LAUNCH(ctx, k_make_keys2, cdiv(m, 256), 256, 0, keys, m, item_bits, rank_bits, k_scratch, ctx->src_a.p);
        size_t tmp = 0;
        CK(cub::DeviceRadixSort::SortPairs(nullptr, tmp, k_scratch, ctx->keys_c.p, ctx->src_a.p, ctx->csc_src.p,
                                           (int64_t)m, 0, std::min(64, item_bits + rank_bits), st));
        ctx->cub_tmp.need(tmp);
        CK(cub::DeviceRadixSort::SortPairs(ctx->cub_tmp.p, tmp, k_scratch, ctx->keys_c.p, ctx->src_a.p, ctx->csc_src.p,
                                           (int64_t)m, 0, std::min(64, item_bits + rank_bits), st));
    }
    const uint64_t* keys2 = ctx->keys_c.p;                   // sorted (item, rank)

    ctx->ifirst.need(TI); ctx->ilast.need(TI); ctx->tend.need(tab); ctx->tloc.need(tab);
    CK(cudaMemsetAsync(ctx->ifirst.p, 0, (size_t)TI * 4, st));
    CK(cudaMemsetAsync(ctx->ilast.p, 0, (size_t)TI * 4, st));
    CK(cudaMemsetAsync(ctx->tstart.p, 0xff, tab * 4, st));
    if (m > 0) LAUNCH(ctx, k_item_groups, cdiv(m, 256), 256, 0, keys2, m, rank_bits, ctx->rank_cluster.p, TI,
                      ctx->ifirst.p, ctx->ilast.p, ctx->tstart.p, ctx->tend.p);
    if (sharded_index) {
        // p(i|C), b and the total were computed from the global item sums before the compaction
    } else if (exact_ok && !ctx->use_ext) {
        ctx->tsum.need(tab);
        LAUNCH(ctx, k_group_sum, cdiv((int64_t)tab, 256), 256, 0, ctx->tstart.p, ctx->tend.p, ctx->csc_src.p, ctx->s_score.p, tab, ctx->tsum.p);
        LAUNCH(ctx, k_item_prob_fast, cdiv(TI, 128), 128, 0, ctx->tsum.p, KC, TI, ctx->counters.p + 1, lambda,
               ctx->isum.p, ctx->iprob.p, ctx->bvec.p, ctx->total.p, ctx->counters.p + 2);
    } else {
        LAUNCH(ctx, k_item_prob, cdiv(TI, 128), 128, 0, ctx->ifirst.p, ctx->ilast.p, ctx->csc_src.p, ctx->s_score.p, TI,
               ctx->counters.p + 1, ctx->use_ext ? ctx->ext_iprob.p : (const double*)nullptr, lambda,
               ctx->isum.p, ctx->iprob.p, ctx->bvec.p, ctx->total.p, ctx->counters.p + 2);
    }
    LAUNCH(ctx, k_cluster_item_count, KC, 256, 0, ctx->tstart.p, TI, ctx->icount.p);
    LAUNCH(ctx, k_cluster_offsets, 1, 32, 0, ctx->icount.p, KC, ctx->item_off.p);

    // per-user work for sharding (replicated-index path; the sharded index did this before compacting)
    ctx->work.need(U); ctx->work_scan.need(U);
    if (!sharded_index) LAUNCH(ctx, k_user_work, cdiv(U, 256), 256, 0, ctx->rowptr.p, ctx->rank_cluster.p, ctx->icount.p, U, ctx->work.p);
    if (ctx->prm.shard_count > 1 && !sharded_index) {
        size_t tmp = 0;
        CK(cub::DeviceScan::InclusiveSum(nullptr, tmp, ctx->work.p, ctx->work_scan.p, U, st));
        ctx->cub_tmp.need(tmp);
        CK(cub::DeviceScan::InclusiveSum(ctx->cub_tmp.p, tmp, ctx->work.p, ctx->work_scan.p, U, st));
    }

    ctx->h_icount.assign(KC, 0); ctx->h_item_off.assign((size_t)KC + 1, 0);
    if (sharded_index) {
        // no sync: the touched clusters hold all their users' ratings, so their item counts are the global ones
        for (int32_t c = 0; c < KC; c++) {
            const int32_t cs = ctx->h_cstart[c], ce = ctx->h_cstart[c + 1];
            const bool touched = ce > cs && cs >= rl_h && ce <= rh_h;
            ctx->h_icount[c] = touched ? h_icount_global[c] : 0;
            ctx->h_item_off[c + 1] = ctx->h_item_off[c] + ctx->h_icount[c];
        }
    } else {
        // sync B: per-cluster item counts, total, flags, (work prefix)
        std::vector<double> h_scan;
        CK(cudaMemcpyAsync(ctx->h_icount.data(), ctx->icount.p, (size_t)KC * 4, cudaMemcpyDeviceToHost, st));
        CK(cudaMemcpyAsync(ctx->h_item_off.data(), ctx->item_off.p, ((size_t)KC + 1) * 4, cudaMemcpyDeviceToHost, st));
        CK(cudaMemcpyAsync(&ctx->h_total, ctx->total.p, sizeof(double), cudaMemcpyDeviceToHost, st));
        CK(cudaMemcpyAsync(&h_bmin_bits, ctx->counters.p + 2, sizeof(h_bmin_bits), cudaMemcpyDeviceToHost, st));
        CK(cudaMemcpyAsync(h_flags, ctx->flags.p, sizeof(h_flags), cudaMemcpyDeviceToHost, st));
        if (world > 1) {
            h_scan.resize(U);
            CK(cudaMemcpyAsync(h_scan.data(), ctx->work_scan.p, (size_t)U * 8, cudaMemcpyDeviceToHost, st));
        }
        CK(cudaStreamSynchronize(st));
        { int rc = check_flags(ctx, h_flags); if (rc != FY_OK) return rc; }
        for (int32_t c = 0; c < KC; c++) h_icount_global[c] = ctx->h_icount[c];
        // shard = contiguous range of user ranks with ~equal estimated work (n_u * I_c)
        if (world > 1) {
            const double tot = h_scan[U - 1];
            std::vector<double> G((size_t)world + 1, 0.0);
            const bool aware = KC <= MAX_PARTITION_CLUSTERS && world <= 64 && shard_beta > 0.0;
            if (aware) {
                std::vector<double> W((size_t)KC, 0.0);
                for (int32_t c = 0; c < KC; c++) {
                    const int32_t a = ctx->h_cstart[c], b = ctx->h_cstart[c + 1];
                    W[c] = (b > a) ? h_scan[b - 1] - (a > 0 ? h_scan[a - 1] : 0.0) : 0.0;
                }
                partition_targets(W.data(), KC, shard_beta, world, G.data());
            }
            for (int r = 1; r < world; r++) {
                const double target = aware ? G[r] : tot * (double)r / (double)world;
                int32_t b = (int32_t)(std::lower_bound(h_scan.begin(), h_scan.end(), target) - h_scan.begin());
                if (aware && b < U) {                     // snap to a cluster boundary the target equals up to rounding (k_shard_bounds)
                    const double tol = 1e-9 * tot;
                    const int32_t c = ctx->h_rank_cluster[b];
                    const int32_t e0 = ctx->h_cstart[c], e1 = ctx->h_cstart[c + 1];
                    const double p0 = e0 > 0 ? h_scan[e0 - 1] : 0.0, p1 = h_scan[e1 - 1];
                    if (std::fabs(p0 - target) <= tol) b = e0;
                    else if (std::fabs(p1 - target) <= tol) b = e1;
                }
                ctx->h_bounds[r] = std::max(ctx->h_bounds[r - 1], b);
            }
            ub = ctx->h_bounds[me]; ue = ctx->h_bounds[me + 1];
        }
    }
    const int32_t n_slots = ctx->h_item_off[KC];
    ctx->shard_begin = ub; ctx->shard_end = ue;

    // ---------------- local item numbering, d, alpha, c(u,j) ----------------
    const int32_t ns1 = std::max(n_slots, 1);
    ctx->c_item.need(ns1); ctx->c_start.need(ns1); ctx->c_len.need(ns1);
    ctx->c_b.need(ns1); ctx->c_alpha.need(ns1);
    ctx->csr_loc.need(m1); ctx->csr_delta.need(m1); ctx->csr_c.need(m1); ctx->csc_lu.need(m1); ctx->csc_delta.need(m1);
    LAUNCH(ctx, k_local_items, KC, 1024, 0, ctx->tstart.p, ctx->tend.p, TI, ctx->item_off.p, ctx->bvec.p,
           ctx->tloc.p, ctx->c_item.p, ctx->c_start.p, ctx->c_len.p, ctx->c_b.p);
    if (m > 0) {
        LAUNCH(ctx, k_delta, cdiv(m, 256), 256, 0, keys, ctx->s_score.p, m, item_bits, ctx->rank_cluster.p, ctx->usum.p,
               ctx->bvec.p, ctx->tloc.p, TI, lambda, ctx->csr_loc.p, ctx->csr_delta.p);
        LAUNCH(ctx, k_csc_fill, cdiv(m, 256), 256, 0, keys2, ctx->csc_src.p, m, rank_bits, ctx->rank_cluster.p, ctx->cstart.p,
               ctx->csr_delta.p, ctx->csc_lu.p, ctx->csc_delta.p);
    }
    ctx->cbound.need((size_t)KC * 3);
    LAUNCH(ctx, k_init_cbound, cdiv(KC, 128), 128, 0, ctx->cbound.p, KC);
    if (n_slots > 0)
        LAUNCH(ctx, k_alpha_cuj, cdiv((int64_t)n_slots * 32, 256), 256, 0, ctx->c_start.p, ctx->c_len.p, ctx->c_b.p, keys2, rank_bits,
               ctx->rank_cluster.p, ctx->cstart.p, ctx->csc_src.p, ctx->csc_delta.p, n_slots, ctx->c_alpha.p, ctx->csr_c.p,
               ctx->cbound.p);
    // processing order of the H-build rows inside a cluster: most raters first
    if (n_slots > 0) {
        ctx->row_keys[0].need(ns1); ctx->row_keys[1].need(ns1); ctx->row_perm.need(ns1); ctx->row_perm_vals.need(ns1);
        LAUNCH(ctx, k_row_perm_keys, cdiv(n_slots, 256), 256, 0, ctx->c_len.p, ctx->item_off.p, KC, n_slots, ctx->row_keys[0].p, ctx->row_perm_vals.p);
        size_t tmp = 0;
        const int end_bit = std::min(64, 32 + bits_for((uint64_t)KC));
        CK(cub::DeviceRadixSort::SortPairs(nullptr, tmp, ctx->row_keys[0].p, ctx->row_keys[1].p, ctx->row_perm_vals.p, ctx->row_perm.p, n_slots, 0, end_bit, st));
        ctx->cub_tmp.need(tmp);
        CK(cub::DeviceRadixSort::SortPairs(ctx->cub_tmp.p, tmp, ctx->row_keys[0].p, ctx->row_keys[1].p, ctx->row_perm_vals.p, ctx->row_perm.p, n_slots, 0, end_bit, st));
    }
    // processing order of the score kernel inside a cluster: most active users first (LPT), so that the 60x-longer
    // CTA of a heavy user never starts at the tail of the grid
    {
        ctx->perm_keys[0].need(U); ctx->perm_keys[1].need(U); ctx->perm_vals.need(U); ctx->perm.need(U);
        LAUNCH(ctx, k_perm_keys, cdiv(U, 256), 256, 0, ctx->rowptr.p, ctx->rank_cluster.p, U, ctx->perm_keys[0].p, ctx->perm_vals.p);
        size_t tmp = 0;
        const int end_bit = std::min(64, 32 + bits_for((uint64_t)KC));
        CK(cub::DeviceRadixSort::SortPairs(nullptr, tmp, ctx->perm_keys[0].p, ctx->perm_keys[1].p, ctx->perm_vals.p, ctx->perm.p, U, 0, end_bit, st));
        ctx->cub_tmp.need(tmp);
        CK(cub::DeviceRadixSort::SortPairs(ctx->cub_tmp.p, tmp, ctx->perm_keys[0].p, ctx->perm_keys[1].p, ctx->perm_vals.p, ctx->perm.p, U, 0, end_bit, st));
    }
    CK(cudaEventRecord(ev_index, st));

    // ---------------- exponent-peel period L from a lower bound on t ----------------
    // t >= (K-1) * b_i * b_j >= b_min^2 (b_min = smallest positive b_i, reduced on the device) and
    // t <= K < 2^31: L factors must stay inside the double range.  b_min = 0 (lambda = 0) -> L = 1,
    // the variant that also handles t = 0 (log 0 = -inf, as Math.log does).
    int L = 1;
    {
        double b_min = 0.0;
        if (h_bmin_bits != ~0ull) std::memcpy(&b_min, &h_bmin_bits, sizeof(double));
        if (b_min > 0.0 && std::isfinite(b_min)) {
            const double neg_log2 = std::max(-2.0 * std::log2(b_min) + 2.0, 40.0);   // |log2 t| <= this
            const int lmax = (int)std::floor(1000.0 / neg_log2);
            L = lmax >= 8 ? 8 : lmax >= 4 ? 4 : lmax >= 2 ? 2 : 1;
        }
    }

    // ---------------- per-cluster scoring ----------------
    const int32_t n_rows = ue - ub;
    int32_t max_ic = 0;
    for (int32_t c = 0; c < KC; c++) max_ic = std::max(max_ic, h_icount_global[c]);   // same stride on every rank
    const int32_t out_stride = std::max(1, std::min(ctx->prm.top_n, max_ic));
    // min(N, I_c) beyond the shared-memory select/sort bound: whole-row stable segmented sort (exact stream)
    const bool big_n = out_stride > TOPN_MAX_SELECT;
    ctx->out_stride = out_stride;
    // With a communicator attached (fy_rm2_comm_init) the dense [row x N] result blocks of all ranks live in one buffer
    // indexed by global user rank and are exchanged in place at the end of the run; otherwise only this shard's rows exist.
    const bool gather_on = ctx->nccl_comm != nullptr && world > 1;
    const int32_t ob = gather_on ? 0 : ub;                   // user rank of output row 0
    const int32_t rows_total = gather_on ? U : n_rows;
    ctx->out_item.need((size_t)std::max(rows_total, 1) * out_stride);
    ctx->out_score.need((size_t)std::max(rows_total, 1) * out_stride);
    ctx->out_count.need((size_t)std::max(rows_total, 1) + 1);
    ctx->out_off.need((size_t)std::max(rows_total, 1) + 1);
    CK(cudaMemsetAsync(ctx->out_count.p, 0, ((size_t)std::max(rows_total, 1) + 1) * 4, st));
    const double log_items = std::log((double)ctx->prm.number_of_items);   // AbstractRM2Reducer.java:328

    // auto mode: approximate stream over the hi-word plane + exact re-score of the candidates
    const bool use_hi = (ctx->prm.score_mode == 0) && !force_exact && L >= 2 && out_stride + 32 <= TOPN_MAX_SELECT;
    int cap = 64; while (cap < out_stride + 32) cap <<= 1;
    ctx->overflow.need(1);
    CK(cudaMemsetAsync(ctx->overflow.p, 0, sizeof(int), st));
    // per-cluster plan of the 4-byte plane: float(H * 2^s) with fp32 math when the range of t allows a
    // peel period >= 2, else the hi-word plane with fp64 math
    struct Plan { int mode; int lf; int sexp; double scale; };
    std::vector<Plan> plan((size_t)KC, Plan{1, 0, 0, 1.0});
    if (use_hi && !ctx->use_ext) {
        std::vector<unsigned long long> hb((size_t)KC * 3);
        CK(cudaMemcpyAsync(hb.data(), ctx->cbound.p, hb.size() * 8, cudaMemcpyDeviceToHost, st));
        CK(cudaMemcpyAsync(h_flags, ctx->flags.p, sizeof(h_flags), cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));      // sync B2: per-cluster bounds on alpha and b (+ duplicate-rating flag of the local sort)
        { int rc = check_flags(ctx, h_flags); if (rc != FY_OK) return rc; }
        for (int32_t c = 0; c < KC; c++) {
            double a_max, b_max, b_min;
            std::memcpy(&a_max, &hb[3 * c], 8); std::memcpy(&b_max, &hb[3 * c + 1], 8);
            if (hb[3 * c + 2] == ~0ull) continue;
            std::memcpy(&b_min, &hb[3 * c + 2], 8);
            const double K = (double)(ctx->h_cstart[c + 1] - ctx->h_cstart[c]);
            if (!(K >= 2.0) || !(b_min > 0.0) || !(a_max > 0.0)) continue;
            // t <= alpha_max (1 + b_max) + b_max (K b_max + alpha_max)   (d <= 1);   t >= (K-1) b_min^2
            const double t_max = a_max * (1.0 + b_max) + b_max * (K * b_max + a_max);
            const double t_min = (K - 1.0) * b_min * b_min;
            const double lo = std::log2(t_min) - 0.5, hi = std::log2(t_max) + 0.5;
            const int sexp = -(int)std::lround(0.5 * (lo + hi));
            const double half = std::max(hi + sexp, -(lo + sexp)) + 1.0;      // |log2(t 2^s)| <= half
            const int lf = (int)std::floor(124.0 / half);
            if (lf >= 2 && std::abs(sexp) < 1000) plan[c] = Plan{2, lf >= 4 ? 4 : 2, sexp, std::ldexp(1.0, sexp)};
        }
    }
    // k_score_f32_tma (rows staged by cp.async.bulk into a shared-memory ring) is kept as a measured alternative, OFF by
    // default: 3.83 vs 3.38 ms per ML-20M cluster, 11.1 vs 8.0 ms per Netflix cluster -- every staged byte is used by exactly
    // one thread, so the ring only adds a second trip through shared memory (2 x 43 GB at 128 B/clk/SM = 2.4 ms by itself)
    const char* tma_env = std::getenv("FY_SCORE_TMA");
    const bool score_tma = (tma_env && std::strcmp(tma_env, "1") == 0);
    const char* spad_env = std::getenv("FY_SCORE_PAD");              // experiment: dynamic shared memory per score CTA (fewer resident CTAs)
    const size_t score_pad = spad_env ? (size_t)std::atoi(spad_env) : 0;
    if (score_pad > 40 * 1024) CK(cudaFuncSetAttribute(k_score_f32<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)score_pad));
    const char* minb_env = std::getenv("FY_SCORE_MINB");
    const int score_minb = minb_env ? std::atoi(minb_env) : 0;
    const char* lpt_env = std::getenv("FY_SCORE_LPT");
    const bool lpt_on = !(lpt_env && std::strcmp(lpt_env, "0") == 0);
    ctx->prof.bytes_per_term = use_hi ? 4.0 : 8.0;
    ctx->prof.score_kernel = 0;
    ctx->prof.exact_rerun = force_exact ? 1 : 0;

    // Three streams: H build of cluster c+1 (latency / DRAM-write bound) and top-N of cluster c-1 run
    // beside the score kernel of cluster c (L1/L2-fabric bound); H and the score matrix are double
    // buffered.  Each stage is timed with events on its own stream.
    // (Stream priorities were tried: measured no gain -- the run sits at the 1 kW power cap, so the
    //  stages do not hide each other; the side streams only fill launch gaps and tails, ~2 %.)
    if (!ctx->stream_g) CK(cudaStreamCreateWithFlags(&ctx->stream_g, cudaStreamNonBlocking));
    if (!ctx->stream_t) CK(cudaStreamCreateWithFlags(&ctx->stream_t, cudaStreamNonBlocking));
    for (cudaEvent_t& e : ctx->sync_ev) if (!e) CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    cudaStream_t sS = st, sG = ctx->stream_g, sT = ctx->stream_t;
    cudaEvent_t* hReady = ctx->sync_ev;       // [3] H[b] built
    cudaEvent_t* hFree = ctx->sync_ev + 3;    // [3] last kernel reading H[b] (score, or the exact re-score) finished
    cudaEvent_t* sReady = ctx->sync_ev + 6;   // [2] scores[b] written
    cudaEvent_t* sFree = ctx->sync_ev + 8;    // [2] top-N finished reading scores[b]
    cudaEvent_t evFork = ctx->sync_ev[10], evJoin = ctx->sync_ev[11];
    const char* hbufs_env = std::getenv("FY_H_BUFS");
    const int n_hbuf = (hbufs_env && std::atoi(hbufs_env) == 3) ? 3 : 2;
    CK(cudaEventRecord(evFork, st));
    CK(cudaStreamWaitEvent(sG, evFork, 0));
    CK(cudaStreamWaitEvent(sT, evFork, 0));
    enum { SEG_GRAM = 0, SEG_SCORE = 1, SEG_TOPN = 2, SEG_REFINE = 3 };
    struct Seg { int kind; size_t e0, e1; };
    std::vector<Seg> segs;
    auto seg_begin = [&](int kind, cudaStream_t strm) { CK(cudaEventRecord(ctx->ev(evi), strm)); segs.push_back(Seg{kind, evi, 0}); evi++; return segs.size() - 1; };
    auto seg_end = [&](size_t k, cudaStream_t strm) { CK(cudaEventRecord(ctx->ev(evi), strm)); segs[k].e1 = evi; evi++; };
    const size_t SCORE_BUF_BYTES = big_n ? ((size_t)1 << 29) : ((size_t)2 << 30);
    // H-build variant: 2 (default) = k_build_H2 (warp per (row, wide column range), rater-sequential, batches of four raters
    // prefetched in registers, fp64 plane written out by cp.async.bulk);
    // 1 = the round-1 flattened-list kernel (kept for A/B: FY_BUILD_H=1).  FY_H2_CFG picks the (range, warps, depth) instance.
    const char* bh_env = std::getenv("FY_BUILD_H");
    const int build_variant = (bh_env && std::strcmp(bh_env, "1") == 0) ? 1 : 2;
    const char* h2_env = std::getenv("FY_H2_CFG");
    const int h2_cfg = h2_env ? std::atoi(h2_env) : 0;
    const char* h2_order_env = std::getenv("FY_H2_ORDER");
    // most-raters-first row order: pays when a row walks many raters (Netflix shape, ~113 raters per row: 4.8 vs 6.5 ms per
    // cluster), costs a little when rows are short and the write locality of the item order matters more (ML-20M shape, ~15
    // raters per row: 2.76 vs 2.64 ms) -- decided by the average raters per row of this run; FY_H2_LPT=0/1 forces it
    const char* h2_lpt_env = std::getenv("FY_H2_LPT");
    const bool h2_lpt = h2_lpt_env ? std::strcmp(h2_lpt_env, "0") != 0 : ((double)m >= 48.0 * (double)std::max(n_slots, 1));
    const char* h2_pad_env = std::getenv("FY_H2_PAD");               // experiment: extra dynamic shared memory per CTA (fewer resident CTAs)
    const size_t h2_pad = h2_pad_env ? (size_t)std::atoi(h2_pad_env) : 0;
    const char* h2_bulk_env = std::getenv("FY_H2_BULK");
    const bool h2_bulk = !(h2_bulk_env && std::strcmp(h2_bulk_env, "0") == 0);
    // Default 1536 columns x 1 warp per CTA: alone, 1024-column ranges are a little faster (2.50 vs 2.56 ms per ML-20M cluster), but inside the job their
    // extra resident warps slow the score kernel running beside them -- 342.5 vs 320.1 ms per ML-20M job, same box, twice
    // (one warp per CTA is 1 % better again than two: 313.3 / 316.5 vs 316.9-320 ms on two boxes)
    int32_t h2_rw = 1536, h2_nw = 1;
    switch (h2_cfg) {
        case 1: h2_rw = 2048; h2_nw = 2; break;
        case 2: h2_rw = 512; h2_nw = 8; break;
        case 3: h2_rw = 1024; h2_nw = 4; break;
        case 4: h2_rw = 768; h2_nw = 4; break;
        case 5: h2_rw = 1024; h2_nw = 2; break;
        case 6: h2_rw = 1280; h2_nw = 2; break;
        case 7: h2_rw = 768; h2_nw = 2; break;
        case 8: h2_rw = 1536; h2_nw = 2; break;
        default: break;
    }
    auto h_geometry = [&](int32_t I_c, int32_t& ld, int32_t& slice_w, int32_t& chunk_w, int32_t& nchunk, int32_t& n_bound) {
        ld = cdiv(I_c, SCOREH_TILE) * SCOREH_TILE;
        if (build_variant == 2) {
            slice_w = h2_rw;
            chunk_w = h2_rw * h2_nw;
            nchunk = cdiv(cdiv(I_c, h2_rw), h2_nw);
            n_bound = cdiv(I_c, h2_rw) + 1;
            return;
        }
        slice_w = H_SLICE;
        chunk_w = H_SLICE * H_WARPS;
        nchunk = cdiv(cdiv(I_c, H_SLICE), H_WARPS);     // CTAs per row
        n_bound = cdiv(I_c, H_SLICE) + 1;               // slice boundaries per user
    };
    auto launch_build = [&](cudaStream_t strm, int32_t I_c, int32_t ld, int32_t nchunk, int32_t n_bound, int32_t slot0,
                            const int32_t* cp, double* Hp, uint32_t* Hhp, int mode, double scale) {
#define FY_H2_LAUNCH_B(RW, NW, PM, BK)                                                                                \
        do {                                                                                                          \
            const size_t smem = (size_t)(NW) * (RW) * 8 + h2_pad;                                                     \
            CK(cudaFuncSetAttribute(k_build_H2<RW, NW, PM, BK>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
            const bool rowmaj = !(h2_order_env && std::strcmp(h2_order_env, "0") == 0);                              \
            LAUNCH_ON(ctx, strm, (k_build_H2<RW, NW, PM, BK>), rowmaj ? dim3((unsigned)I_c * (unsigned)nchunk) : dim3(I_c, nchunk), (NW) * 32, smem, I_c, ld, n_bound - 1, slot0,  \
                      ctx->c_start.p, ctx->c_len.p, ctx->c_b.p, ctx->c_alpha.p, ctx->csc_lu.p, ctx->csc_delta.p, cp,  \
                      ctx->csr_loc.p, ctx->csr_delta.p, Hp, Hhp, scale, rowmaj ? nchunk : 0,                          \
                      h2_lpt ? ctx->row_perm.p : (const int32_t*)nullptr);                                            \
        } while (0)
#define FY_H2_LAUNCH_PM(RW, NW, PM)                                                                                   \
        do { if (h2_bulk) FY_H2_LAUNCH_B(RW, NW, PM, true); else FY_H2_LAUNCH_B(RW, NW, PM, false); } while (0)
#define FY_H2_LAUNCH(RW, NW)                                                                                          \
        do {                                                                                                          \
            if (!Hhp) FY_H2_LAUNCH_PM(RW, NW, 0);                                                                     \
            else if (mode == 2) FY_H2_LAUNCH_PM(RW, NW, 2);                                                           \
            else FY_H2_LAUNCH_PM(RW, NW, 1);                                                                          \
        } while (0)
        if (build_variant == 2) {
            switch (h2_cfg) {
                case 1: FY_H2_LAUNCH(2048, 2); break;
                case 2: FY_H2_LAUNCH(512, 8); break;
                case 3: FY_H2_LAUNCH(1024, 4); break;
                case 4: FY_H2_LAUNCH(768, 4); break;
                case 5: FY_H2_LAUNCH(1024, 2); break;
                case 6: FY_H2_LAUNCH(1280, 2); break;
                case 7: FY_H2_LAUNCH(768, 2); break;
                case 8: FY_H2_LAUNCH(1536, 2); break;
                default: FY_H2_LAUNCH(1536, 1); break;
            }
            return;
        }
#undef FY_H2_LAUNCH
#undef FY_H2_LAUNCH_PM
#undef FY_H2_LAUNCH_B
        const size_t smem = (size_t)H_SLICE * H_WARPS * sizeof(double);      // 8 warps x H_SLICE doubles
        if (smem > 48 * 1024) CK(cudaFuncSetAttribute(k_build_H, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        LAUNCH_ON(ctx, strm, k_build_H, dim3(I_c, nchunk), H_THREADS, smem, I_c, ld, n_bound - 1, slot0,
                  ctx->c_start.p, ctx->c_len.p, ctx->c_b.p, ctx->c_alpha.p, ctx->csc_lu.p, ctx->csc_delta.p,
                  cp, ctx->csr_loc.p, ctx->csr_delta.p, Hp, Hhp, mode, scale);
    };
    int n_touched_total = 0;
    {   // size the per-cluster buffers once (growing them inside the loop would synchronise the device)
        size_t need_h = 0, need_cp = 0, need_sc = 0, need_us = 0;
        int n_touched = 0, n_batches = 0;
        for (int32_t c = 0; c < KC; c++) {
            const int32_t cs = ctx->h_cstart[c], ce = ctx->h_cstart[c + 1];
            int32_t r0 = std::max(cs, ub), r1 = std::min(ce, ue);
            const int32_t I_c = ctx->h_icount[c];
            if (!ctx->owner_rank.empty()) { r0 = std::max(r0, ctx->owner_rank[c]); r1 = std::min(r1, ctx->owner_rank[c] + 1); }
            if (r1 <= r0 || I_c <= 0) continue;
            int32_t ld, slice_w, chunk_w, nchunk, n_bound;
            h_geometry(I_c, ld, slice_w, chunk_w, nchunk, n_bound);
            const size_t batch = std::max<size_t>(1, std::min<size_t>((size_t)(r1 - r0), SCORE_BUF_BYTES / ((size_t)ld * 8)));
            need_h = std::max(need_h, (size_t)I_c * ld);
            need_cp = std::max(need_cp, (size_t)(ce - cs) * n_bound);
            need_sc = std::max(need_sc, batch * ld);
            need_us = std::max(need_us, batch * 3);
            n_touched++; n_batches += cdiv(r1 - r0, (int64_t)batch);
        }
        n_touched_total = n_touched;
        for (int b = 0; b < 3; b++) {
            const bool h_on = b == 0 || (b < n_hbuf && n_touched > b);          // a single cluster needs one H
            const bool s_on = b == 0 || (b == 1 && n_batches >= 2);
            if (h_on) { ctx->H[b].need(need_h); ctx->chunk_ptr2[b].need(need_cp); if (use_hi) ctx->Hh[b].need(need_h); }
            if (s_on) {
                ctx->scores[b].need(need_sc); ctx->ustat[b].need(need_us);
                if (use_hi) { ctx->cand[b].need(need_us / 3 * cap); ctx->cand_score[b].need(need_us / 3 * cap); ctx->cand_cnt[b].need(need_us / 3); }
            }
            if (big_n && b == 0) {
                ctx->sort_keys[0].need(need_sc); ctx->sort_keys[1].need(need_sc); ctx->sort_idx[0].need(need_sc); ctx->sort_idx[1].need(need_sc);
                ctx->seg_off.need(need_us / 3 + 2);
                if (need_sc >= ((size_t)1 << 31)) return ctx->fail(FY_E_UNSUPPORTED, "score batch too large for the segmented sort");
            }
        }
    }
    const int h_cycle = std::min(n_hbuf, std::max(1, n_touched_total));
    int hb = 0, sb = 0;
    bool h_used[3] = {false, false, false}, s_used[2] = {false, false};
    for (int32_t c = 0; c < KC; c++) {
        const int32_t cs = ctx->h_cstart[c], ce = ctx->h_cstart[c + 1];
        int32_t r0 = std::max(cs, ub), r1 = std::min(ce, ue);
        // neighbour-list mode: the cluster is {u} + N(u) and only its owner u is scored (fy_rm2_run_neighbours)
        if (!ctx->owner_rank.empty()) { r0 = std::max(r0, ctx->owner_rank[c]); r1 = std::min(r1, ctx->owner_rank[c] + 1); }
        if (r1 <= r0) continue;
        const int32_t K_c = ce - cs, I_c = ctx->h_icount[c], slot0 = ctx->h_item_off[c];
        if (I_c <= 0) continue;
        int32_t ld, slice_w, chunk_w, nchunk, n_bound;
        h_geometry(I_c, ld, slice_w, chunk_w, nchunk, n_bound);
        ctx->prof.gram_bytes += (double)I_c * ld * 8.0;
        // ---- stream G: H[hb] ----
        if (h_used[hb]) CK(cudaStreamWaitEvent(sG, hFree[hb], 0));
        {
            const size_t k = seg_begin(SEG_GRAM, sG);
            LAUNCH_ON(ctx, sG, k_chunk_ptr, cdiv((int64_t)K_c * n_bound, 256), 256, 0, cs, K_c, n_bound, slice_w,
                      ctx->rowptr.p, ctx->csr_loc.p, ctx->chunk_ptr2[hb].p);
            launch_build(sG, I_c, ld, nchunk, n_bound, slot0, ctx->chunk_ptr2[hb].p, ctx->H[hb].p,
                         use_hi ? ctx->Hh[hb].p : (uint32_t*)nullptr, plan[c].mode, plan[c].scale);
            seg_end(k, sG);
        }
        CK(cudaEventRecord(hReady[hb], sG));
        h_used[hb] = true;
        // ---- stream S: scores; stream T: top-N ----
        CK(cudaStreamWaitEvent(sS, hReady[hb], 0));
        const double log_K = std::log((double)K_c);                        // :329
        const int32_t batch = (int32_t)std::max<size_t>(1, std::min<size_t>((size_t)(r1 - r0), SCORE_BUF_BYTES / ((size_t)ld * 8)));
        int P2 = 1; while (P2 < std::min(std::min(ctx->prm.top_n, I_c), (int32_t)TOPN_MAX_SELECT)) P2 <<= 1;
        const size_t topn_smem = (size_t)P2 * 12;
        if (!big_n && topn_smem > 36 * 1024) CK(cudaFuncSetAttribute(k_topn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)topn_smem));
        if (use_hi && (size_t)cap * 12 > 40 * 1024) CK(cudaFuncSetAttribute(k_refine_sort, cudaFuncAttributeMaxDynamicSharedMemorySize, cap * 12));
        for (int32_t b0 = r0; b0 < r1; b0 += batch) {
            const int32_t nb = std::min(batch, r1 - b0);
            const dim3 grid(nb, use_hi ? ld / SCOREH_TILE : ld / SCORE_TILE);
            // the LPT order is a permutation of the cluster's ranks: usable when this batch is the whole cluster
            const int32_t* lpt = (lpt_on && b0 == cs && nb == ce - cs) ? ctx->perm.p : (const int32_t*)nullptr;
            if (s_used[sb]) CK(cudaStreamWaitEvent(sS, sFree[sb], 0));
            LAUNCH_ON(ctx, sS, k_init_ustat, cdiv(nb, 256), 256, 0, ctx->ustat[sb].p, nb);
            {
                const size_t k = seg_begin(SEG_SCORE, sS);
                if (use_hi) ctx->prof.score_kernel = std::max(ctx->prof.score_kernel, plan[c].mode);
                if (use_hi && plan[c].mode == 2 && score_tma) {
                    const dim3 g2(nb, ld / SCOREH_TILE);
                    if (plan[c].lf >= 4) {
                        CK(cudaFuncSetAttribute(k_score_f32_tma<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SCORE_TMA_SMEM));
                        LAUNCH_ON(ctx, sS, k_score_f32_tma<4>, g2, SCORE_TMA_THREADS, SCORE_TMA_SMEM, ctx->Hh[hb].p, I_c, ld, b0, slot0, ctx->rowptr.p, ctx->csr_loc.p, ctx->csr_c.p, ctx->c_b.p, plan[c].scale, plan[c].sexp, log_items, log_K, ctx->scores[sb].p, ctx->ustat[sb].p, lpt);
                    } else {
                        CK(cudaFuncSetAttribute(k_score_f32_tma<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SCORE_TMA_SMEM));
                        LAUNCH_ON(ctx, sS, k_score_f32_tma<2>, g2, SCORE_TMA_THREADS, SCORE_TMA_SMEM, ctx->Hh[hb].p, I_c, ld, b0, slot0, ctx->rowptr.p, ctx->csr_loc.p, ctx->csr_c.p, ctx->c_b.p, plan[c].scale, plan[c].sexp, log_items, log_K, ctx->scores[sb].p, ctx->ustat[sb].p, lpt);
                    }
                } else if (use_hi && plan[c].mode == 2 && score_minb > 1 && plan[c].lf >= 4) {
                    // experiment (FY_SCORE_MINB): the same kernel compiled for a minimum residency of 6 / 10 / 12 CTAs per SM
#define FY_SCORE_MB(MB) LAUNCH_ON(ctx, sS, (k_score_f32<4, MB>), grid, SCORE_THREADS, 0, ctx->Hh[hb].p, I_c, ld, b0, slot0, ctx->rowptr.p, ctx->csr_loc.p, ctx->csr_c.p, ctx->c_b.p, plan[c].scale, plan[c].sexp, log_items, log_K, ctx->scores[sb].p, ctx->ustat[sb].p, lpt)
                    if (score_minb >= 12) FY_SCORE_MB(12); else if (score_minb >= 10) FY_SCORE_MB(10); else FY_SCORE_MB(6);
#undef FY_SCORE_MB
                } else if (use_hi && plan[c].mode == 2) {
                    if (plan[c].lf >= 4)
                        LAUNCH_ON(ctx, sS, k_score_f32<4>, grid, SCORE_THREADS, score_pad, ctx->Hh[hb].p, I_c, ld, b0, slot0, ctx->rowptr.p, ctx->csr_loc.p, ctx->csr_c.p, ctx->c_b.p, plan[c].scale, plan[c].sexp, log_items, log_K, ctx->scores[sb].p, ctx->ustat[sb].p, lpt);
                    else
                        LAUNCH_ON(ctx, sS, k_score_f32<2>, grid, SCORE_THREADS, 0, ctx->Hh[hb].p, I_c, ld, b0, slot0, ctx->rowptr.p, ctx->csr_loc.p, ctx->csr_c.p, ctx->c_b.p, plan[c].scale, plan[c].sexp, log_items, log_K, ctx->scores[sb].p, ctx->ustat[sb].p, lpt);
                } else if (use_hi) {
                    switch (L) {
                        case 8: LAUNCH_ON(ctx, sS, k_score_hi<8>, grid, SCORE_THREADS, 0, ctx->Hh[hb].p, I_c, ld, b0, slot0, ctx->rowptr.p, ctx->csr_loc.p, ctx->csr_c.p, ctx->c_b.p, log_items, log_K, ctx->scores[sb].p, ctx->ustat[sb].p); break;
                        case 4: LAUNCH_ON(ctx, sS, k_score_hi<4>, grid, SCORE_THREADS, 0, ctx->Hh[hb].p, I_c, ld, b0, slot0, ctx->rowptr.p, ctx->csr_loc.p, ctx->csr_c.p, ctx->c_b.p, log_items, log_K, ctx->scores[sb].p, ctx->ustat[sb].p); break;
                        default: LAUNCH_ON(ctx, sS, k_score_hi<2>, grid, SCORE_THREADS, 0, ctx->Hh[hb].p, I_c, ld, b0, slot0, ctx->rowptr.p, ctx->csr_loc.p, ctx->csr_c.p, ctx->c_b.p, log_items, log_K, ctx->scores[sb].p, ctx->ustat[sb].p); break;
                    }
                } else {
                    switch (L) {
                        case 8: launch_score<8>(ctx, sS, grid, ctx->H[hb].p, I_c, ld, b0, slot0, log_items, log_K, ctx->scores[sb].p, ctx->ustat[sb].p); break;
                        case 4: launch_score<4>(ctx, sS, grid, ctx->H[hb].p, I_c, ld, b0, slot0, log_items, log_K, ctx->scores[sb].p, ctx->ustat[sb].p); break;
                        case 2: launch_score<2>(ctx, sS, grid, ctx->H[hb].p, I_c, ld, b0, slot0, log_items, log_K, ctx->scores[sb].p, ctx->ustat[sb].p); break;
                        default: launch_score<1>(ctx, sS, grid, ctx->H[hb].p, I_c, ld, b0, slot0, log_items, log_K, ctx->scores[sb].p, ctx->ustat[sb].p); break;
                    }
                }
                seg_end(k, sS);
            }
            ctx->prof.score_launches++;
            CK(cudaEventRecord(sReady[sb], sS));
            s_used[sb] = true;
            CK(cudaStreamWaitEvent(sT, sReady[sb], 0));
            if (big_n) {
                const size_t k = seg_begin(SEG_TOPN, sT);
                const int64_t tot = (int64_t)nb * ld;
                LAUNCH_ON(ctx, sT, k_sort_prepare, cdiv(tot, 256), 256, 0, ctx->scores[sb].p, I_c, ld, tot, ctx->sort_keys[0].p, ctx->sort_idx[0].p);
                LAUNCH_ON(ctx, sT, k_seg_offsets, cdiv(nb + 1, 256), 256, 0, ctx->seg_off.p, nb, ld);
                size_t tmp = 0;
                CK(cub::DeviceSegmentedSort::StableSortPairs(nullptr, tmp, ctx->sort_keys[0].p, ctx->sort_keys[1].p, ctx->sort_idx[0].p,
                                                             ctx->sort_idx[1].p, tot, nb, ctx->seg_off.p, ctx->seg_off.p + 1, sT));
                if (tmp > ctx->sort_tmp.cap) { CK(cudaStreamSynchronize(sT)); ctx->sort_tmp.need(tmp); }
                CK(cub::DeviceSegmentedSort::StableSortPairs(ctx->sort_tmp.p, tmp, ctx->sort_keys[0].p, ctx->sort_keys[1].p, ctx->sort_idx[0].p,
                                                             ctx->sort_idx[1].p, tot, nb, ctx->seg_off.p, ctx->seg_off.p + 1, sT));
                LAUNCH_ON(ctx, sT, k_emit_sorted, nb, 256, 0, ctx->sort_keys[1].p, ctx->sort_idx[1].p, ctx->ustat[sb].p, ld, b0, ob, slot0,
                          ctx->prm.top_n, out_stride, ctx->prm.filter_users, ctx->split, ctx->n_splits, ctx->rank_userid.p,
                          ctx->c_item.p, ctx->out_item.p, ctx->out_score.p, ctx->out_count.p);
                seg_end(k, sT);
            } else {
                const size_t k = seg_begin(SEG_TOPN, sT);
                LAUNCH_ON(ctx, sT, k_topn, nb, TOPN_THREADS, topn_smem, ctx->scores[sb].p, ctx->ustat[sb].p, I_c, ld, b0, slot0,
                          ctx->prm.top_n, out_stride, ctx->prm.filter_users, ctx->split, ctx->n_splits, ctx->rank_userid.p,
                          ctx->c_item.p, b0 - ob, ctx->out_item.p, ctx->out_score.p, ctx->out_count.p,
                          ctx->rowptr.p, cap, plan[c].mode == 2 ? 2.5e-7 : 4.8e-7, use_hi ? ctx->cand[sb].p : (int32_t*)nullptr,
                          use_hi ? ctx->cand_cnt[sb].p : (int32_t*)nullptr, ctx->overflow.p);
                seg_end(k, sT);
            }
            if (use_hi) {
                const size_t k = seg_begin(SEG_REFINE, sT);
                LAUNCH_ON(ctx, sT, k_refine_score, dim3(nb, cdiv(cap, REFINE_THREADS / 32)), REFINE_THREADS, 0, ctx->H[hb].p, ld, b0,
                          slot0, ctx->rowptr.p, ctx->csr_loc.p, ctx->csr_c.p, ctx->c_b.p, log_items, log_K, cap, ctx->cand[sb].p,
                          ctx->cand_cnt[sb].p, ctx->cand_score[sb].p);
                LAUNCH_ON(ctx, sT, k_refine_sort, nb, REFINE_THREADS, (size_t)cap * 12, b0, ob, slot0, ctx->c_item.p, cap,
                          ctx->cand[sb].p, ctx->cand_cnt[sb].p, ctx->cand_score[sb].p, out_stride, ctx->out_item.p,
                          ctx->out_score.p, ctx->out_count.p);
                seg_end(k, sT);
            }
            CK(cudaEventRecord(sFree[sb], sT));
            if (ctx->scores[1].cap) sb ^= 1;
        }
        CK(cudaEventRecord(hFree[hb], use_hi ? sT : sS));     // the exact re-score reads H too
        hb = (hb + 1) % h_cycle;
        ctx->prof.clusters_touched++;
    }
    // join
    CK(cudaEventRecord(evJoin, sG)); CK(cudaStreamWaitEvent(st, evJoin, 0));
    CK(cudaEventRecord(evFork, sT)); CK(cudaStreamWaitEvent(st, evFork, 0));

    // sync C1: did a candidate list overflow?  (decided before the exchange: every rank enters the collective exactly once)
    int h_overflow = 0;
    CK(cudaMemcpyAsync(&h_overflow, ctx->overflow.p, sizeof(int), cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(h_flags, ctx->flags.p, sizeof(h_flags), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    { int rc = check_flags(ctx, h_flags); if (rc != FY_OK) return rc; }
    if (use_hi && h_overflow > 0) return 1;      // redo in exact mode (rare)

    // ---------------- exchange: every rank ends with the top-N blocks of all users (north_star: "NCCL ... to gather the
    // final top-N lists"); one grouped in-place broadcast per rank and array, issued from here on the run's stream ----------------
    cudaEvent_t ev_gather0 = ctx->ev(evi++), ev_gather1 = ctx->ev(evi++);
    CK(cudaEventRecord(ev_gather0, st));
    if (gather_on) {
        int rc = nccl_exchange(ctx, out_stride, st);
        if (rc != FY_OK) return rc;
    }
    CK(cudaEventRecord(ev_gather1, st));

    // ---------------- pack ----------------
    ctx->counters.need(8); ctx->run_terms.need(1);
    CK(cudaMemsetAsync(ctx->counters.p + 4, 0, sizeof(unsigned long long), st));
    CK(cudaMemsetAsync(ctx->run_terms.p, 0, sizeof(double), st));
    const size_t bound_out = (size_t)std::max(rows_total, 1) * out_stride;
    ctx->p_user.need(bound_out); ctx->p_item.need(bound_out); ctx->p_cluster.need(bound_out);
    ctx->p_s64.need(bound_out); ctx->p_s32.need(bound_out);
    ctx->row_user.need((size_t)std::max(rows_total, 1)); ctx->row_cluster.need((size_t)std::max(rows_total, 1));
    if (rows_total > 0) {
        size_t tmp = 0;
        ctx->out_cnt64.need((size_t)rows_total + 1);
        LAUNCH(ctx, k_widen_counts, cdiv(rows_total + 1, 256), 256, 0, ctx->out_count.p, rows_total + 1, ctx->out_cnt64.p);
        CK(cub::DeviceScan::ExclusiveSum(nullptr, tmp, ctx->out_cnt64.p, ctx->out_off.p, rows_total + 1, st));
        ctx->cub_tmp.need(tmp);
        CK(cub::DeviceScan::ExclusiveSum(ctx->cub_tmp.p, tmp, ctx->out_cnt64.p, ctx->out_off.p, rows_total + 1, st));
        LAUNCH(ctx, k_pack, rows_total, 128, 0, ctx->out_item.p, ctx->out_score.p, ctx->out_count.p, ctx->out_off.p, out_stride,
               ob, rows_total, ctx->rank_userid.p, ctx->rank_cluster.p, ctx->p_user.p, ctx->p_item.p, ctx->p_s64.p, ctx->p_s32.p,
               ctx->p_cluster.p, ctx->row_user.p, ctx->row_cluster.p);
    }
    if (n_rows > 0)      // users scored and log-terms of THIS shard, reduced on the device
        LAUNCH(ctx, k_run_totals, cdiv(n_rows, 256), 256, 0, ctx->out_count.p + (ub - ob), ctx->work.p + ub, n_rows,
               ctx->counters.p + 4, ctx->run_terms.p);
    CK(cudaEventRecord(ev_end, st));
    int64_t total_out = 0;
    unsigned long long h_users = 0;
    double terms = 0.0;
    if (rows_total > 0) CK(cudaMemcpyAsync(&total_out, ctx->out_off.p + rows_total, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(&h_users, ctx->counters.p + 4, sizeof(h_users), cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(&terms, ctx->run_terms.p, sizeof(double), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));   // sync C2: number of result triples, users scored, work figures
    ctx->n_results = total_out;
    ctx->n_result_rows = rows_total;
    ctx->users_scored = (int64_t)h_users;
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, ev_gather0, ev_gather1)); ctx->prof.ms_gather = ms;
    CK(cudaEventElapsedTime(&ms, ev_start, ev_end)); ctx->prof.ms_total = ms;
    CK(cudaEventElapsedTime(&ms, ev_start, ev_index)); ctx->prof.ms_index = ms;
    for (const Seg& g : segs) {
        CK(cudaEventElapsedTime(&ms, ctx->ev(g.e0), ctx->ev(g.e1)));
        if (g.kind == SEG_GRAM) ctx->prof.ms_gram += ms;
        else if (g.kind == SEG_SCORE) ctx->prof.ms_score += ms;
        else if (g.kind == SEG_TOPN) ctx->prof.ms_topn += ms;
        else if (g.kind == SEG_REFINE) ctx->prof.ms_refine += ms;
    }
    ctx->prof.log_terms = terms;
    ctx->prof.score_bytes = ctx->prof.bytes_per_term * terms;
    ctx->prof.users_scored = ctx->users_scored;
    ctx->prof.kernel_launches = ctx->launches;
    ctx->have_results = true;
    return FY_OK;
}

static int run_pipeline(fy_rm2_ctx* ctx) {
    int rc = run_pipeline_impl(ctx, false);
    if (rc == 1) rc = run_pipeline_impl(ctx, true);
    return rc;
}

extern "C" int fy_rm2_run(fy_rm2_ctx* ctx) {
    if (!ctx) return FY_E_ARG;
    ctx->view = nullptr;
    if (!ctx->kids.empty()) {
        ctx->have_results = false;
        const int rc = fan_out(ctx, [&](fy_rm2_ctx* k, size_t) { return fy_rm2_run(k); });
        if (rc != FY_OK) return rc;
        // shards are contiguous ranges of the (cluster, user id) order: concatenating them in shard order keeps it
        ctx->kid_off.assign(ctx->kids.size() + 1, 0);
        ctx->users_scored = 0; ctx->n_result_rows = 0;
        fy_rm2_profile agg{};
        for (size_t i = 0; i < ctx->kids.size(); i++) {
            const fy_rm2_ctx* k = ctx->kids[i];
            ctx->kid_off[i + 1] = ctx->kid_off[i] + k->n_results;
            ctx->users_scored += k->users_scored;
            ctx->n_result_rows += k->n_result_rows;
            const fy_rm2_profile& q = k->prof;
            agg.ms_total = std::max(agg.ms_total, q.ms_total); agg.ms_index = std::max(agg.ms_index, q.ms_index);
            agg.ms_gram = std::max(agg.ms_gram, q.ms_gram); agg.ms_score = std::max(agg.ms_score, q.ms_score);
            agg.ms_topn = std::max(agg.ms_topn, q.ms_topn); agg.ms_refine = std::max(agg.ms_refine, q.ms_refine);
            agg.log_terms += q.log_terms; agg.score_bytes += q.score_bytes; agg.gram_bytes += q.gram_bytes;
            agg.users_scored += q.users_scored; agg.kernel_launches += q.kernel_launches;
            agg.clusters_touched += q.clusters_touched; agg.score_launches += q.score_launches;
            agg.bytes_per_term = q.bytes_per_term; agg.exact_rerun = std::max(agg.exact_rerun, q.exact_rerun);
            agg.score_kernel = std::max(agg.score_kernel, q.score_kernel);
        }
        ctx->prof = agg;
        ctx->n_results = ctx->kid_off.back();
        ctx->have_results = true;
        return FY_OK;
    }
    return guarded(ctx, [&]() { return run_pipeline(ctx); });
}

extern "C" int32_t fy_rm2_max_item(const fy_rm2_ctx* ctx) { return !ctx ? -1 : (ctx->kids.empty() ? ctx->max_item : ctx->kids[0]->max_item); }

extern "C" int64_t fy_rm2_user_count(const fy_rm2_ctx* ctx) {
    if (!ctx) return -1;
    if (!ctx->kids.empty()) return ctx->kids[0]->have_clustering ? ctx->kids[0]->n_users : -1;
    return ctx->have_clustering || ctx->have_results ? ctx->n_users : -1;
}

extern "C" int fy_rm2_stats(fy_rm2_ctx* ctx, double* user_sum, double* item_prob, double* total) {
    if (!ctx) return FY_E_ARG;
    if (!ctx->have_results) return ctx->fail(FY_E_STATE, "fy_rm2_stats needs a successful fy_rm2_run");
    if (!ctx->kids.empty()) {            // the statistics are global: every device holds the same ones
        const int rc = fy_rm2_stats(ctx->kids[0], user_sum, item_prob, total);
        return rc == FY_OK ? rc : ctx->fail(rc, "%s", ctx->kids[0]->err);
    }
    return guarded(ctx, [&]() {
        CK(cudaSetDevice(ctx->prm.device));
        if (user_sum) {
            std::vector<double> tmp((size_t)ctx->n_users);
            CK(cudaMemcpyAsync(tmp.data(), ctx->usum.p, (size_t)ctx->n_users * 8, cudaMemcpyDeviceToHost, ctx->stream));
            CK(cudaStreamSynchronize(ctx->stream));
            for (int32_t k = 0; k < ctx->n_users; k++) user_sum[k] = tmp[ctx->h_input_rank[k]];
        }
        if (item_prob) {
            CK(cudaMemcpyAsync(item_prob, ctx->iprob.p, ((size_t)ctx->max_item + 1) * 8, cudaMemcpyDeviceToHost, ctx->stream));
            CK(cudaStreamSynchronize(ctx->stream));
        }
        if (total) *total = ctx->h_total;
        return (int)FY_OK;
    });
}

extern "C" int64_t fy_rm2_result_count(const fy_rm2_ctx* ctx) { if (ctx && ctx->view) ctx = ctx->view; return (ctx && ctx->have_results) ? ctx->n_results : -1; }
extern "C" int64_t fy_rm2_users_scored(const fy_rm2_ctx* ctx) { if (ctx && ctx->view) ctx = ctx->view; return (ctx && ctx->have_results) ? ctx->users_scored : -1; }

extern "C" int fy_rm2_results(fy_rm2_ctx* ctx, int32_t* user, int32_t* item, double* score64, float* score32, int32_t* cluster) {
    if (!ctx) return FY_E_ARG;
    if (ctx->view) { const int rc = fy_rm2_results(ctx->view, user, item, score64, score32, cluster); return rc == FY_OK ? rc : ctx->fail(rc, "%s", ctx->view->err); }
    if (!ctx->have_results) return ctx->fail(FY_E_STATE, "fy_rm2_results needs a successful fy_rm2_run");
    if (!ctx->kids.empty())              // every device copies its slice straight into the caller's buffers, in parallel
        return fan_out(ctx, [&](fy_rm2_ctx* k, size_t i) {
            const int64_t o = ctx->kid_off[i];
            return fy_rm2_results(k, user ? user + o : nullptr, item ? item + o : nullptr, score64 ? score64 + o : nullptr,
                                  score32 ? score32 + o : nullptr, cluster ? cluster + o : nullptr);
        });
    return guarded(ctx, [&]() {
        CK(cudaSetDevice(ctx->prm.device));
        const size_t n = (size_t)ctx->n_results;
        cudaStream_t st = ctx->stream;
        if (n) {
            if (user) CK(cudaMemcpyAsync(user, ctx->p_user.p, n * 4, cudaMemcpyDeviceToHost, st));
            if (item) CK(cudaMemcpyAsync(item, ctx->p_item.p, n * 4, cudaMemcpyDeviceToHost, st));
            if (score64) CK(cudaMemcpyAsync(score64, ctx->p_s64.p, n * 8, cudaMemcpyDeviceToHost, st));
            if (score32) CK(cudaMemcpyAsync(score32, ctx->p_s32.p, n * 4, cudaMemcpyDeviceToHost, st));
            if (cluster) CK(cudaMemcpyAsync(cluster, ctx->p_cluster.p, n * 4, cudaMemcpyDeviceToHost, st));
        }
        CK(cudaStreamSynchronize(st));
        return (int)FY_OK;
    });
}

extern "C" int fy_rm2_results_device(fy_rm2_ctx* ctx, const int32_t** user, const int32_t** item, const double** score64,
                                     const float** score32, const int32_t** cluster) {
    if (!ctx) return FY_E_ARG;
    if (!ctx->have_results) return ctx->fail(FY_E_STATE, "fy_rm2_results_device needs a successful fy_rm2_run");
    if (!ctx->kids.empty()) return ctx->fail(FY_E_UNSUPPORTED, "fy_rm2_results_device on an n_gpus > 1 context (results live on several devices)");
    if (user) *user = ctx->p_user.p;
    if (item) *item = ctx->p_item.p;
    if (score64) *score64 = ctx->p_s64.p;
    if (score32) *score32 = ctx->p_s32.p;
    if (cluster) *cluster = ctx->p_cluster.p;
    return FY_OK;
}

extern "C" int64_t fy_rm2_result_row_count(const fy_rm2_ctx* ctx) { if (ctx && ctx->view) ctx = ctx->view; return (ctx && ctx->have_results) ? ctx->n_result_rows : -1; }

extern "C" int fy_rm2_result_rows(fy_rm2_ctx* ctx, int32_t* user, int32_t* cluster, int32_t* count) {
    if (!ctx) return FY_E_ARG;
    if (ctx->view) { const int rc = fy_rm2_result_rows(ctx->view, user, cluster, count); return rc == FY_OK ? rc : ctx->fail(rc, "%s", ctx->view->err); }
    if (!ctx->have_results) return ctx->fail(FY_E_STATE, "fy_rm2_result_rows needs a successful fy_rm2_run");
    if (!ctx->kids.empty()) {
        std::vector<int64_t> ro(ctx->kids.size() + 1, 0);
        for (size_t i = 0; i < ctx->kids.size(); i++) ro[i + 1] = ro[i] + ctx->kids[i]->n_result_rows;
        return fan_out(ctx, [&](fy_rm2_ctx* k, size_t i) {
            return fy_rm2_result_rows(k, user ? user + ro[i] : nullptr, cluster ? cluster + ro[i] : nullptr, count ? count + ro[i] : nullptr);
        });
    }
    return guarded(ctx, [&]() {
        CK(cudaSetDevice(ctx->prm.device));
        const size_t n = (size_t)ctx->n_result_rows;
        cudaStream_t st = ctx->stream;
        if (n) {
            if (user) CK(cudaMemcpyAsync(user, ctx->row_user.p, n * 4, cudaMemcpyDeviceToHost, st));
            if (cluster) CK(cudaMemcpyAsync(cluster, ctx->row_cluster.p, n * 4, cudaMemcpyDeviceToHost, st));
            if (count) CK(cudaMemcpyAsync(count, ctx->out_count.p, n * 4, cudaMemcpyDeviceToHost, st));
        }
        CK(cudaStreamSynchronize(st));
        return (int)FY_OK;
    });
}

// Roofline probe (see k_probe_plane_read): GB/s the L2 -> SM path delivers for the score kernel's access pattern.
extern "C" int fy_rm2_probe_plane_read(fy_rm2_ctx* ctx, int32_t n_rows, int32_t n_users, int32_t rows_per_user, int32_t reps,
                                       double* gb_per_s, double* ms_per_launch) {
    if (!ctx || n_rows <= 0 || n_users <= 0 || rows_per_user <= 0 || reps <= 0) return FY_E_ARG;
    if (!ctx->kids.empty()) return ctx->fail(FY_E_UNSUPPORTED, "not available on an n_gpus > 1 context");
    return guarded(ctx, [&]() {
        using namespace fy;
        CK(cudaSetDevice(ctx->prm.device));
        cudaStream_t st = ctx->stream;
        const int32_t ld = cdiv(n_rows, SCOREH_TILE) * SCOREH_TILE;
        DBuf<uint32_t> plane, sink;
        plane.need((size_t)n_rows * ld); sink.need(4);
        CK(cudaMemsetAsync(plane.p, 0x3f, (size_t)n_rows * ld * 4, st));
        const int32_t rpu = cdiv(rows_per_user, 8) * 8;
        const dim3 grid(n_users, ld / SCOREH_TILE);
        LAUNCH(ctx, k_probe_plane_read, grid, SCORE_THREADS, 0, plane.p, n_rows, ld, rpu, sink.p);      // warm-up
        cudaEvent_t e0 = ctx->ev(0), e1 = ctx->ev(1);
        CK(cudaEventRecord(e0, st));
        for (int r = 0; r < reps; r++) LAUNCH(ctx, k_probe_plane_read, grid, SCORE_THREADS, 0, plane.p, n_rows, ld, rpu, sink.p);
        CK(cudaEventRecord(e1, st));
        CK(cudaStreamSynchronize(st));
        float ms = 0;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        const double bytes = (double)n_users * ld * 4.0 * rpu;
        if (ms_per_launch) *ms_per_launch = ms / reps;
        if (gb_per_s) *gb_per_s = bytes * reps / (ms * 1e-3) / 1e9;
        return (int)FY_OK;
    });
}

extern "C" int fy_rm2_get_profile(const fy_rm2_ctx* ctx, fy_rm2_profile* out) {
    if (!ctx || !out) return FY_E_ARG;
    if (ctx->view) ctx = ctx->view;
    *out = ctx->prof;
    return FY_OK;
}

// ---------------------------------------------------------------------------------------------
// f3 / north_star part 1: scoring over EXPLICIT neighbour lists.  buildRecommendations takes `int[] neighbours`
// (M/rm/AbstractRM2Reducer.java:321-323,342-346); the reference only ever passes "the cluster minus u" (:215-216).
// Here user u with list N(u) is scored exactly as the reducer would score it in a group made of u and N(u):
//   K = |N(u)| + 1 (:143,:329), items = everything u or a neighbour rated (:164-174), sum over v in N(u) (:342-346),
//   p(i|C) and the user sums are the GLOBAL statistics of the rating matrix (jobs RM2-1/2).
// Implementation: the host expands every listed user into such a group (its members' ratings under fresh virtual user
// ids, members in ascending real id so that the summation order is the reducer's canonical one), a child context runs
// the ordinary per-cluster pipeline on the groups with the global statistics passed in (the fine seam's mechanism) and
// scores only each group's owner; ids are mapped back.  No reference counterpart exists beyond the signature:
// checked against the neighbour mode of the CPU restatement (tests/), which reproduces the 507 goldens when N(u) = cluster(u) \ {u}.
// ---------------------------------------------------------------------------------------------
extern "C" int fy_rm2_run_neighbours(fy_rm2_ctx* ctx, const int32_t* user, const int32_t* neighbour, int32_t k, int64_t n_listed) {
    if (!ctx) return FY_E_ARG;
    if (!user || !neighbour || k <= 0 || n_listed <= 0) return ctx->fail(FY_E_ARG, "fy_rm2_run_neighbours: bad argument");
    if (!ctx->kids.empty() || ctx->prm.shard_count > 1) return ctx->fail(FY_E_UNSUPPORTED, "fy_rm2_run_neighbours on a sharded / n_gpus > 1 context");
    if (!ctx->have_ratings) return ctx->fail(FY_E_STATE, "fy_rm2_run_neighbours needs fy_rm2_set_ratings first");
    ctx->view = nullptr;
    return guarded(ctx, [&]() {
        CK(cudaSetDevice(ctx->prm.device));
        const int64_t nnz = ctx->nnz;
        // ---- the ratings back on the host, positive ones only, by (user id, item id) ----
        std::vector<int32_t> hu((size_t)nnz), hi((size_t)nnz);
        std::vector<float> hs((size_t)nnz);
        CK(cudaMemcpyAsync(hu.data(), ctx->in_user.p, (size_t)nnz * 4, cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaMemcpyAsync(hi.data(), ctx->in_item.p, (size_t)nnz * 4, cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaMemcpyAsync(hs.data(), ctx->in_score.p, (size_t)nnz * 4, cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        std::vector<int64_t> ord;
        ord.reserve((size_t)nnz);
        for (int64_t e = 0; e < nnz; e++) if (hs[e] > 0.0f) ord.push_back(e);            // ScoreByClusterHDFSMapper.java:39-40
        std::sort(ord.begin(), ord.end(), [&](int64_t a, int64_t b) { return hu[a] != hu[b] ? hu[a] < hu[b] : hi[a] < hi[b]; });
        for (size_t x = 1; x < ord.size(); x++)
            if (hu[ord[x]] == hu[ord[x - 1]] && hi[ord[x]] == hi[ord[x - 1]]) return ctx->fail(FY_E_DUPLICATE_RATING, "the same (user,item) pair is rated twice");
        // per user: [begin, end) in ord, and the user sum (ascending item order, DoubleSumReducer.java:31-42)
        std::vector<int32_t> uid;
        std::vector<int64_t> ubeg;
        std::vector<double> usum;
        const int32_t TI = ctx->max_item + 1;
        std::vector<double> isum((size_t)TI, 0.0), iprob((size_t)TI, 0.0);
        long long counter = 0;
        for (size_t x = 0; x < ord.size();) {
            size_t y = x;
            double s = 0.0;
            while (y < ord.size() && hu[ord[y]] == hu[ord[x]]) { s += (double)hs[ord[y]]; isum[hi[ord[y]]] += (double)hs[ord[y]]; y++; }
            uid.push_back(hu[ord[x]]); ubeg.push_back((int64_t)x); usum.push_back(s);
            counter += (long long)s * 100;                                                  // DoubleSumAndCountReducer.java:41
            x = y;
        }
        ubeg.push_back((int64_t)ord.size());
        const double total = (double)counter / 100.0;                                        // RM2Job.java:95
        for (int32_t i = 0; i < TI; i++) iprob[i] = isum[i] > 0.0 ? isum[i] / total : 0.0;    // DoubleSumAndDividerReducer.java:44
        auto find_user = [&](int32_t id) -> int64_t {
            const auto it = std::lower_bound(uid.begin(), uid.end(), id);
            return (it != uid.end() && *it == id) ? (int64_t)(it - uid.begin()) : -1;
        };
        // ---- expansion: one virtual cluster per listed user ----
        std::vector<int32_t> xu, xi, vuser, vcluster, vsize((size_t)n_listed), owner_vid((size_t)n_listed), vreal;
        std::vector<float> xs;
        std::vector<double> vsum;
        std::vector<int32_t> members;
        int64_t next_vid = 1;
        for (int64_t q = 0; q < n_listed; q++) {
            members.clear();
            members.push_back(user[q]);
            for (int32_t t = 0; t < k; t++) { const int32_t v = neighbour[q * k + t]; if (v >= 0 && v != user[q]) members.push_back(v); }
            std::sort(members.begin(), members.end());
            members.erase(std::unique(members.begin(), members.end()), members.end());
            vsize[q] = (int32_t)members.size();
            for (const int32_t mreal : members) {
                const int64_t ux = find_user(mreal);
                if (ux < 0) return ctx->fail(FY_E_USER_WITHOUT_RATING, "user %d (listed, or a neighbour of user %d) has no positive rating", mreal, user[q]);
                if (next_vid >= 0x7fffffff) return ctx->fail(FY_E_UNSUPPORTED, "too many (user, neighbour) pairs for one call: split the user list");
                const int32_t vid = (int32_t)next_vid++;
                if (mreal == user[q]) owner_vid[q] = vid;
                vuser.push_back(vid); vcluster.push_back((int32_t)q); vreal.push_back(mreal); vsum.push_back(usum[ux]);
                for (int64_t x = ubeg[ux]; x < ubeg[ux + 1]; x++) { xu.push_back(vid); xi.push_back(hi[ord[x]]); xs.push_back(hs[ord[x]]); }
            }
        }
        if ((int64_t)xu.size() > 0x7fffffffll - 1024 || n_listed * (int64_t)TI > (1ll << 28))
            return ctx->fail(FY_E_UNSUPPORTED, "neighbour-list job too large for one call (%lld expanded ratings, %lld listed users x %d item ids): split the user list",
                             (long long)xu.size(), (long long)n_listed, TI);
        // ---- child context: the ordinary pipeline on the groups, global statistics passed in, owners only ----
        if (!ctx->nbr_child) {
            fy_rm2_params kp = ctx->prm;
            kp.filter_users = 0; kp.shard_rank = 0; kp.shard_count = 1; kp.n_gpus = 0;
            const int rc = fy_rm2_create(&ctx->nbr_child, &kp);
            if (rc != FY_OK) return ctx->fail(rc, "could not create the neighbour-list context");
        }
        fy_rm2_ctx* ch = ctx->nbr_child;
        int rc = upload_ratings(ch, xu.data(), xi.data(), xs.data(), (int64_t)xu.size());
        if (rc == FY_OK) rc = upload_clustering(ch, vuser.data(), vcluster.data(), (int64_t)vuser.size(), vsize.data(), (int32_t)n_listed, true);
        if (rc != FY_OK) return ctx->fail(rc, "%s", ch->err);
        const int32_t V = (int32_t)vuser.size();
        std::vector<double> us_rank((size_t)V);
        for (int32_t j = 0; j < V; j++) us_rank[ch->h_input_rank[j]] = vsum[j];
        ch->owner_rank.assign((size_t)n_listed, 0);
        std::vector<int32_t> rank_of_vid((size_t)V + 1, 0);
        for (int32_t j = 0; j < V; j++) rank_of_vid[vuser[j]] = ch->h_input_rank[j];
        for (int64_t q = 0; q < n_listed; q++) ch->owner_rank[q] = rank_of_vid[owner_vid[q]];
        ch->ext_usum.need(V); ch->ext_iprob.need((size_t)ch->max_item + 1);
        CK(cudaMemcpyAsync(ch->ext_usum.p, us_rank.data(), (size_t)V * 8, cudaMemcpyHostToDevice, ch->stream));
        CK(cudaMemcpyAsync(ch->ext_iprob.p, iprob.data(), ((size_t)ch->max_item + 1) * 8, cudaMemcpyHostToDevice, ch->stream));
        CK(cudaStreamSynchronize(ch->stream));
        ch->use_ext = true; ch->split = 0; ch->n_splits = 1;
        rc = run_pipeline(ch);
        ch->use_ext = false;
        ch->owner_rank.clear();
        ch->have_ratings = false; ch->have_clustering = false;
        if (rc != FY_OK) return ctx->fail(rc, "%s", ch->err);
        // ---- virtual ids back to the callers' user ids (cluster = position of the user in the call) ----
        if (ch->n_results > 0) {
            std::vector<int32_t> pu((size_t)ch->n_results);
            CK(cudaMemcpyAsync(pu.data(), ch->p_user.p, (size_t)ch->n_results * 4, cudaMemcpyDeviceToHost, ch->stream));
            CK(cudaStreamSynchronize(ch->stream));
            for (int32_t& v : pu) v = vreal[(size_t)v - 1];
            CK(cudaMemcpyAsync(ch->p_user.p, pu.data(), (size_t)ch->n_results * 4, cudaMemcpyHostToDevice, ch->stream));
        }
        if (ch->n_result_rows > 0) {
            std::vector<int32_t> ru((size_t)ch->n_result_rows);
            CK(cudaMemcpyAsync(ru.data(), ch->row_user.p, (size_t)ch->n_result_rows * 4, cudaMemcpyDeviceToHost, ch->stream));
            CK(cudaStreamSynchronize(ch->stream));
            for (int32_t& v : ru) v = vreal[(size_t)v - 1];
            CK(cudaMemcpyAsync(ch->row_user.p, ru.data(), (size_t)ch->n_result_rows * 4, cudaMemcpyHostToDevice, ch->stream));
        }
        CK(cudaStreamSynchronize(ch->stream));
        ctx->view = ch;
        return (int)FY_OK;
    });
}

// ---------------------------------------------------------------------------------------------
// fine seam: one AbstractRM2Reducer.reduce() group (M/rm/AbstractRM2Reducer.java:129-233)
// ---------------------------------------------------------------------------------------------
extern "C" int fy_rm2_score_group(fy_rm2_ctx* ctx, int32_t cluster_id, int32_t split, int32_t n_splits,
                                  const int32_t* group_user, const double* group_user_sum, int32_t n_group_users,
                                  const int32_t* r_user, const int32_t* r_item, const float* r_score, int64_t nnz,
                                  const double* item_prob, int32_t max_item) {
    if (!ctx) return FY_E_ARG;
    if (!group_user || !group_user_sum || !r_user || !r_item || !r_score || !item_prob || n_group_users <= 0 || nnz <= 0 ||
        n_splits <= 0 || split < 0 || split >= n_splits || cluster_id < 0 || max_item < 0)
        return ctx->fail(FY_E_ARG, "fy_rm2_score_group: bad argument");
    if (ctx->prm.shard_count > 1 || !ctx->kids.empty()) return ctx->fail(FY_E_UNSUPPORTED, "fy_rm2_score_group on a sharded / n_gpus > 1 context");
    return guarded(ctx, [&]() {
        int rc = upload_ratings(ctx, r_user, r_item, r_score, nnz);
        if (rc != FY_OK) return rc;
        if (ctx->max_item > max_item) return ctx->fail(FY_E_ARG, "p(%d|C) not found (item id beyond item_prob)", ctx->max_item);  // :294-298
        // a single cluster holding exactly the group's users; ids keep their cluster id for the sink
        std::vector<int32_t> cl((size_t)n_group_users, 0), sz(1, n_group_users);
        rc = upload_clustering(ctx, group_user, cl.data(), n_group_users, sz.data(), 1, true);
        if (rc != FY_OK) return rc;
        std::vector<double> us((size_t)n_group_users);
        for (int32_t k = 0; k < n_group_users; k++) us[ctx->h_input_rank[k]] = group_user_sum[k];
        ctx->ext_usum.need(n_group_users); ctx->ext_iprob.need((size_t)ctx->max_item + 1);
        CK(cudaMemcpyAsync(ctx->ext_usum.p, us.data(), (size_t)n_group_users * 8, cudaMemcpyHostToDevice, ctx->stream));
        CK(cudaMemcpyAsync(ctx->ext_iprob.p, item_prob, ((size_t)ctx->max_item + 1) * 8, cudaMemcpyHostToDevice, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        ctx->use_ext = true; ctx->split = split; ctx->n_splits = n_splits;
        rc = run_pipeline(ctx);
        ctx->use_ext = false; ctx->split = 0; ctx->n_splits = 1;
        ctx->have_ratings = false; ctx->have_clustering = false;   // the group replaced the coarse inputs
        if (rc != FY_OK) return rc;
        // report the caller's cluster id at the sink (writePreference(..., cluster), :364-365)
        if (ctx->n_results > 0) {
            std::vector<int32_t> cid((size_t)ctx->n_results, cluster_id);
            CK(cudaMemcpyAsync(ctx->p_cluster.p, cid.data(), (size_t)ctx->n_results * 4, cudaMemcpyHostToDevice, ctx->stream));
            CK(cudaMemcpyAsync(ctx->row_cluster.p, cid.data(), (size_t)std::min<int64_t>(ctx->n_result_rows, ctx->n_results) * 4, cudaMemcpyHostToDevice, ctx->stream));
            CK(cudaStreamSynchronize(ctx->stream));
        }
        return (int)FY_OK;
    });
}

// ---------------------------------------------------------------------------------------------
// config 3: item-item co-occurrence (GEMM kernel + launcher in cooc_tcgen05.cu)
// ---------------------------------------------------------------------------------------------
extern "C" int fyi_cooc_gemm_launch(const uint8_t* Bt, int n_items, int k_pad, int32_t* C, int ldc, void* stream,
                                    char* err, size_t errlen);

extern "C" int fy_cooc_counts(fy_rm2_ctx* ctx, int32_t n_user_ids, int32_t n_items, int32_t* counts_out, double* ms_gemm_out) {
    if (!ctx) return FY_E_ARG;
    if (!ctx->kids.empty()) return ctx->fail(FY_E_UNSUPPORTED, "not available on an n_gpus > 1 context");
    if (!ctx->have_ratings) return ctx->fail(FY_E_STATE, "fy_cooc_counts needs fy_rm2_set_ratings first");
    if (n_user_ids <= 0 || n_items <= ctx->max_item) return ctx->fail(FY_E_ARG, "n_items must exceed the largest rated item id (%d)", ctx->max_item);
    return guarded(ctx, [&]() {
        using namespace fy;
        CK(cudaSetDevice(ctx->prm.device));
        cudaStream_t st = ctx->stream;
        const int32_t k_pad = cdiv(n_user_ids, 128) * 128, ldc = cdiv(n_items, 256) * 256;
        ctx->cooc_bt.need((size_t)n_items * k_pad);
        ctx->cooc_counts.need((size_t)n_items * ldc);
        ctx->flags.need(DF_COUNT);
        CK(cudaMemsetAsync(ctx->flags.p, 0, sizeof(int) * DF_COUNT, st));
        CK(cudaMemsetAsync(ctx->cooc_bt.p, 0, (size_t)n_items * k_pad, st));
        LAUNCH(ctx, k_binarise, cdiv(ctx->nnz, 256), 256, 0, ctx->in_user.p, ctx->in_item.p, ctx->in_score.p, ctx->nnz,
               n_user_ids, n_items, k_pad, ctx->cooc_bt.p, ctx->flags.p);
        cudaEvent_t e0 = ctx->ev(0), e1 = ctx->ev(1);
        CK(cudaEventRecord(e0, st));
        const int rc = fyi_cooc_gemm_launch(ctx->cooc_bt.p, n_items, k_pad, ctx->cooc_counts.p, ldc, (void*)st, ctx->err, sizeof(ctx->err));
        if (rc != 0) return rc;
        ctx->launches++;
        CK(cudaEventRecord(e1, st));
        int h_flags[DF_COUNT];
        CK(cudaMemcpyAsync(h_flags, ctx->flags.p, sizeof(h_flags), cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        if (h_flags[DF_BAD_ITEM]) return ctx->fail(FY_E_ARG, "user or item id outside [0, n_user_ids) x [0, n_items)");
        float ms = 0;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        if (ms_gemm_out) *ms_gemm_out = ms;
        ctx->cooc_items = n_items; ctx->cooc_ldc = ldc;
        if (counts_out)
            CK(cudaMemcpy2D(counts_out, (size_t)n_items * 4, ctx->cooc_counts.p, (size_t)ldc * 4, (size_t)n_items * 4, n_items, cudaMemcpyDeviceToHost));
        return (int)FY_OK;
    });
}

extern "C" int fyi_gemm_u8_nt(const uint8_t* A, int a_rows, const uint8_t* B, int b_rows, int k_pad, int32_t* C, int ldc,
                              int symmetric, void* stream, char* err, size_t errlen);

// ---------------------------------------------------------------------------------------------
// a9 / f3: kNN neighbourhood provider.  user-user co-occurrence counts (B B^T on the binarised matrix,
// the same int8 tcgen05 GEMM, in blocks of rows because U^2 counts do not fit) and per-user top-k
// neighbours (k_topn; self excluded, ties by ascending user id).  The reference has NO such step
// (its neighbourhood is the user's cluster): checked against an integer CPU restatement
// only; not wired into the RM2 scoring.
// ---------------------------------------------------------------------------------------------
extern "C" int fy_knn_neighbours(fy_rm2_ctx* ctx, int32_t n_user_ids, int32_t n_items, int32_t k,
                                 int32_t* neighbour_out, int32_t* count_out, int32_t* n_out, double* ms_gemm_out) {
    if (!ctx) return FY_E_ARG;
    if (!ctx->kids.empty()) return ctx->fail(FY_E_UNSUPPORTED, "not available on an n_gpus > 1 context");
    if (!ctx->have_ratings) return ctx->fail(FY_E_STATE, "fy_knn_neighbours needs fy_rm2_set_ratings first");
    if (!neighbour_out || n_user_ids <= 0 || n_items <= ctx->max_item) return ctx->fail(FY_E_ARG, "bad argument (n_items must exceed the largest rated item id %d)", ctx->max_item);
    if (k <= 0 || k > fy::TOPN_MAX_SELECT) return ctx->fail(FY_E_UNSUPPORTED, "k outside [1, %d]", fy::TOPN_MAX_SELECT);
    return guarded(ctx, [&]() {
        using namespace fy;
        CK(cudaSetDevice(ctx->prm.device));
        cudaStream_t st = ctx->stream;
        const int32_t U = n_user_ids;
        const int32_t k_pad = cdiv(n_items, 128) * 128, ldc = cdiv(U, 256) * 256, ld = cdiv(U, SCORE_TILE) * SCORE_TILE;
        const int32_t R = (int32_t)std::min<int64_t>(cdiv(U, 128) * 128, 4096);          // rows of C per GEMM
        const int32_t batch = (int32_t)std::max<size_t>(1, std::min<size_t>((size_t)R, ((size_t)1 << 30) / ((size_t)ld * 8)));
        ctx->cooc_bt.need((size_t)U * k_pad);
        ctx->cooc_counts.need((size_t)R * ldc);
        ctx->scores[0].need((size_t)batch * ld); ctx->ustat[0].need((size_t)batch * 3);
        ctx->cooc_iota.need(U); ctx->cooc_zero.need(U);
        ctx->cooc_out_item.need((size_t)U * k); ctx->cooc_out_score.need((size_t)U * k); ctx->cooc_out_cnt.need(U);
        ctx->flags.need(DF_COUNT);
        CK(cudaMemsetAsync(ctx->flags.p, 0, sizeof(int) * DF_COUNT, st));
        CK(cudaMemsetAsync(ctx->cooc_bt.p, 0, (size_t)U * k_pad, st));
        // Bu[user][item]: k_binarise with the roles of the two ids swapped
        LAUNCH(ctx, k_binarise, cdiv(ctx->nnz, 256), 256, 0, ctx->in_item.p, ctx->in_user.p, ctx->in_score.p, ctx->nnz,
               n_items, U, k_pad, ctx->cooc_bt.p, ctx->flags.p);
        LAUNCH(ctx, k_iota, cdiv(U, 256), 256, 0, ctx->cooc_iota.p, U);
        CK(cudaMemsetAsync(ctx->cooc_zero.p, 0, (size_t)U * 4, st));
        CK(cudaMemsetAsync(ctx->cooc_out_cnt.p, 0, (size_t)U * 4, st));
        int P2 = 1; while (P2 < std::min(k, U)) P2 <<= 1;
        const size_t topn_smem = (size_t)P2 * 12;
        if (topn_smem > 36 * 1024) CK(cudaFuncSetAttribute(k_topn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)topn_smem));
        double ms_gemm = 0.0;
        std::vector<std::pair<cudaEvent_t, cudaEvent_t>> evs;
        size_t evi = 0;
        for (int32_t r0 = 0; r0 < U; r0 += R) {
            const int32_t rows = std::min(R, U - r0);
            cudaEvent_t e0 = ctx->ev(evi++), e1 = ctx->ev(evi++);
            CK(cudaEventRecord(e0, st));
            const int rc = fyi_gemm_u8_nt(ctx->cooc_bt.p + (size_t)r0 * k_pad, rows, ctx->cooc_bt.p, U, k_pad, ctx->cooc_counts.p, ldc, 0,
                                          (void*)st, ctx->err, sizeof(ctx->err));
            if (rc != 0) return rc;
            ctx->launches++;
            CK(cudaEventRecord(e1, st));
            evs.emplace_back(e0, e1);
            for (int32_t b0 = 0; b0 < rows; b0 += batch) {
                const int32_t nb = std::min(batch, rows - b0);
                LAUNCH(ctx, k_init_ustat, cdiv(nb, 256), 256, 0, ctx->ustat[0].p, nb);
                LAUNCH(ctx, k_cooc_scores, dim3(nb, ld / SCORE_TILE), SCORE_THREADS, 0, ctx->cooc_counts.p, U, ldc, ld, b0, r0,
                       ctx->scores[0].p, ctx->ustat[0].p);
                LAUNCH(ctx, k_topn, nb, TOPN_THREADS, topn_smem, ctx->scores[0].p, ctx->ustat[0].p, U, ld, 0, 0, k, k, 0, 0, 1,
                       ctx->cooc_zero.p, ctx->cooc_iota.p, r0 + b0, ctx->cooc_out_item.p, ctx->cooc_out_score.p, ctx->cooc_out_cnt.p,
                       (const int32_t*)nullptr, 0, 0.0, (int32_t*)nullptr, (int32_t*)nullptr, (int*)nullptr);
            }
        }
        int h_flags[DF_COUNT];
        std::vector<double> sc((size_t)U * k);
        std::vector<int32_t> cnt((size_t)U);
        CK(cudaMemcpyAsync(h_flags, ctx->flags.p, sizeof(h_flags), cudaMemcpyDeviceToHost, st));
        CK(cudaMemcpyAsync(neighbour_out, ctx->cooc_out_item.p, (size_t)U * k * 4, cudaMemcpyDeviceToHost, st));
        CK(cudaMemcpyAsync(sc.data(), ctx->cooc_out_score.p, (size_t)U * k * 8, cudaMemcpyDeviceToHost, st));
        CK(cudaMemcpyAsync(cnt.data(), ctx->cooc_out_cnt.p, (size_t)U * 4, cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        if (h_flags[DF_BAD_ITEM]) return ctx->fail(FY_E_ARG, "user or item id outside [0, n_user_ids) x [0, n_items)");
        for (auto& e : evs) { float ms = 0; CK(cudaEventElapsedTime(&ms, e.first, e.second)); ms_gemm += ms; }
        if (ms_gemm_out) *ms_gemm_out = ms_gemm;
        for (int32_t r = 0; r < U; r++) {
            if (n_out) n_out[r] = cnt[r];
            for (int32_t t = 0; t < k; t++) {
                const size_t o = (size_t)r * k + t;
                if (t < cnt[r]) { if (count_out) count_out[o] = (int32_t)sc[o]; }
                else { neighbour_out[o] = -1; if (count_out) count_out[o] = 0; }
            }
        }
        ctx->cooc_items = 0;       // the count matrix no longer holds item-item counts
        return (int)FY_OK;
    });
}

extern "C" int fy_cooc_topk(fy_rm2_ctx* ctx, int32_t k, int32_t* item_out, int32_t* count_out, int32_t* n_out) {
    if (!ctx) return FY_E_ARG;
    if (ctx->cooc_items <= 0) return ctx->fail(FY_E_STATE, "fy_cooc_topk needs fy_cooc_counts first");
    if (k <= 0 || k > fy::TOPN_MAX_SELECT) return ctx->fail(FY_E_UNSUPPORTED, "k outside [1, %d]", fy::TOPN_MAX_SELECT);
    return guarded(ctx, [&]() {
        using namespace fy;
        CK(cudaSetDevice(ctx->prm.device));
        cudaStream_t st = ctx->stream;
        const int32_t n = ctx->cooc_items, ldc = ctx->cooc_ldc, ld = cdiv(n, SCORE_TILE) * SCORE_TILE;
        const int32_t batch = (int32_t)std::max<size_t>(1, std::min<size_t>((size_t)n, ((size_t)1 << 30) / ((size_t)ld * 8)));
        ctx->scores[0].need((size_t)batch * ld); ctx->ustat[0].need((size_t)batch * 3);
        ctx->cooc_iota.need(n); ctx->cooc_zero.need(n);
        ctx->cooc_out_item.need((size_t)n * k); ctx->cooc_out_score.need((size_t)n * k); ctx->cooc_out_cnt.need(n);
        LAUNCH(ctx, k_iota, cdiv(n, 256), 256, 0, ctx->cooc_iota.p, n);
        CK(cudaMemsetAsync(ctx->cooc_zero.p, 0, (size_t)n * 4, st));
        CK(cudaMemsetAsync(ctx->cooc_out_cnt.p, 0, (size_t)n * 4, st));
        int P2 = 1; while (P2 < std::min(k, n)) P2 <<= 1;
        const size_t topn_smem = (size_t)P2 * 12;
        if (topn_smem > 36 * 1024) CK(cudaFuncSetAttribute(k_topn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)topn_smem));
        for (int32_t r0 = 0; r0 < n; r0 += batch) {
            const int32_t nb = std::min(batch, n - r0);
            LAUNCH(ctx, k_init_ustat, cdiv(nb, 256), 256, 0, ctx->ustat[0].p, nb);
            LAUNCH(ctx, k_cooc_scores, dim3(nb, ld / SCORE_TILE), SCORE_THREADS, 0, ctx->cooc_counts.p, n, ldc, ld, r0, 0,
                   ctx->scores[0].p, ctx->ustat[0].p);
            LAUNCH(ctx, k_topn, nb, TOPN_THREADS, topn_smem, ctx->scores[0].p, ctx->ustat[0].p, n, ld, 0, 0, k, k, 0, 0, 1,
                   ctx->cooc_zero.p, ctx->cooc_iota.p, r0, ctx->cooc_out_item.p, ctx->cooc_out_score.p, ctx->cooc_out_cnt.p,
                   (const int32_t*)nullptr, 0, 0.0, (int32_t*)nullptr, (int32_t*)nullptr, (int*)nullptr);
        }
        std::vector<double> sc((size_t)n * k);
        std::vector<int32_t> cnt((size_t)n);
        CK(cudaMemcpyAsync(item_out, ctx->cooc_out_item.p, (size_t)n * k * 4, cudaMemcpyDeviceToHost, st));
        CK(cudaMemcpyAsync(sc.data(), ctx->cooc_out_score.p, (size_t)n * k * 8, cudaMemcpyDeviceToHost, st));
        CK(cudaMemcpyAsync(cnt.data(), ctx->cooc_out_cnt.p, (size_t)n * 4, cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        for (int32_t r = 0; r < n; r++) {
            if (n_out) n_out[r] = cnt[r];
            for (int32_t t = 0; t < k; t++) {
                const size_t o = (size_t)r * k + t;
                if (t < cnt[r]) { if (count_out) count_out[o] = (int32_t)sc[o]; }
                else { item_out[o] = -1; if (count_out) count_out[o] = 0; }
            }
        }
        return (int)FY_OK;
    });
}
