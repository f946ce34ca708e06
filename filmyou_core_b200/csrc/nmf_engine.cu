// nmf_engine.cu -- NMF / PPC clustering on one B200 (C ABI: include/filmyou_nmf.h).
//
// The reference runs one iteration as 9 chained MapReduce jobs (M/nmf/hcomputation/ComputeHJob.java:74-101,
// M/nmf/wcomputation/ComputeWJob.java:72-98).  Here the ratings are indexed once (CSR by user, CSC by item,
// two cub radix sorts) and stay resident; one iteration is 8 small kernels captured in a CUDA graph:
//
//   k_join_partial / k_join_combine   X_H = A^T W   and   X_W = A H     (gathers of k-double factor rows; the factor
//                                     matrices are L2-resident, so the bound is the L2 -> SM path, not HBM)
//   k_cross_partial / k_cross_combine C_W = W^T W   and   C_H = H^T H   (k x k, fp64)
//   k_update<H|W>                     y = C f,  PPC / NMF multiplicative update, fused
//
// Both updates read the OLD H and W (AbstractNMFDriver.java:118-124 builds hJob and wJob on the same paths).
// Every sum has a fixed order (the combiner structure of fy_nmf_params), multiplies and adds are rounded
// separately (__dmul_rn / __dadd_rn: Java never contracts), so results are bit-identical run to run.
#include "../../include/filmyou_nmf.h"

#include <cuda_runtime.h>
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include <algorithm>
#include <cfloat>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <new>
#include <vector>

namespace {

struct NmfCudaFail { cudaError_t err; const char* what; int line; };
#define NCK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) throw NmfCudaFail{e_, #call, __LINE__}; } while (0)

template <class T>
struct NBuf {
    T* p = nullptr;
    size_t cap = 0;
    void need(size_t n) {
        if (n <= cap) return;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        cudaError_t e = cudaMalloc((void**)&p, std::max<size_t>(n, 1) * sizeof(T));
        if (e != cudaSuccess) { p = nullptr; throw NmfCudaFail{e, "cudaMalloc", __LINE__}; }
        cap = n;
    }
    ~NBuf() { if (p) cudaFree(p); }
};

inline int ncdiv(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

constexpr double NMF_EPS = 1e-12;            // M/nmf/MatrixComputationJob.java:41
constexpr int CROSS_THREADS = 256;
constexpr int CROSS_EPT = 8;                 // C entries per thread of k_cross_partial
constexpr int CROSS_TILE = 16;               // factor rows staged in shared memory at a time
constexpr int MAX_K = 512;

enum { NF_BAD_ID = 0, NF_COUNT = 4 };

// ---------------------------------------------------------------------------------------------
// index
// ---------------------------------------------------------------------------------------------
// rating -> (row << 32 | col) keys for both orders; score <= 0 gets the all-ones key (sorts last)
__global__ void k_nmf_keys(const int32_t* __restrict__ user, const int32_t* __restrict__ item, const float* __restrict__ score,
                           int64_t nnz, int32_t id_base, int32_t n_users, int32_t n_items,
                           uint64_t* __restrict__ key_u, uint64_t* __restrict__ key_i,
                           unsigned long long* __restrict__ n_valid, int* __restrict__ flags) {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int ok = 0;
    if (e < nnz) {
        uint64_t ku = ~0ull, ki = ~0ull;
        if (score[e] > 0.0f) {                                   // VectorByItemHDFSMapper.java:40-42
            const int64_t u = (int64_t)user[e] - id_base, i = (int64_t)item[e] - id_base;
            if (u < 0 || u >= n_users || i < 0 || i >= n_items) atomicOr(&flags[NF_BAD_ID], 1);
            else { ku = ((uint64_t)u << 32) | (uint64_t)i; ki = ((uint64_t)i << 32) | (uint64_t)u; ok = 1; }
        }
        key_u[e] = ku; key_i[e] = ki;
    }
    const int cnt = __syncthreads_count(ok);
    if (threadIdx.x == 0 && cnt) atomicAdd(n_valid, (unsigned long long)cnt);
}

// rowptr[r] = first sorted position whose row >= r; rows without entries are reported (smallest id)
__global__ void k_nmf_rowptr(const uint64_t* __restrict__ keys, int32_t m, int32_t n_rows, int32_t* __restrict__ rowptr) {
    const int32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r > n_rows) return;
    int32_t a = 0, b = m;
    while (a < b) { const int32_t mid = (a + b) >> 1; if ((int32_t)(keys[mid] >> 32) < r) a = mid + 1; else b = mid; }
    rowptr[r] = a;
}

__global__ void k_nmf_segcount(const int32_t* __restrict__ rowptr, int32_t n_rows, int32_t combine_len,
                               int32_t* __restrict__ nseg, int* __restrict__ first_empty) {
    const int32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_rows) return;
    const int32_t len = rowptr[r + 1] - rowptr[r];
    if (len == 0) atomicMin(first_empty, r);
    nseg[r] = (combine_len > 0) ? (len + combine_len - 1) / combine_len : (len > 0 ? 1 : 0);
}

__global__ void k_nmf_segfill(const int32_t* __restrict__ seg_off, int32_t n_rows, int32_t* __restrict__ seg_row) {
    const int32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_rows) return;
    for (int32_t s = seg_off[r]; s < seg_off[r + 1]; s++) seg_row[s] = r;
}

// ---------------------------------------------------------------------------------------------
// X[r] = sum over the ratings of row r of F[col] * score   (VectorSumReducer as combiner + reducer)
// blockDim = (TX, R): one thread row per combiner segment, TX threads stride the k columns.
// ---------------------------------------------------------------------------------------------
template <int CPT>   // factor columns per thread: 2 (one 16-byte gather per rating) when k is even, else 1
__global__ void __launch_bounds__(256)
k_join_partial(const int32_t* __restrict__ seg_row, const int32_t* __restrict__ seg_off, int32_t n_seg,
               const int32_t* __restrict__ rowptr, const uint64_t* __restrict__ keys, const float* __restrict__ score,
               const double* __restrict__ F, int32_t k, int32_t combine_len,
               double* __restrict__ part, double* __restrict__ X) {
    const int32_t s = blockIdx.x * blockDim.y + threadIdx.y;
    if (s >= n_seg) return;
    const int32_t r = seg_row[s];
    const int32_t s0 = seg_off[r], ns = seg_off[r + 1] - s0;
    const int32_t b = rowptr[r] + (s - s0) * combine_len;
    const int32_t e = (combine_len > 0) ? min(b + combine_len, rowptr[r + 1]) : rowptr[r + 1];
    double* __restrict__ dst = (ns == 1) ? X + (size_t)r * k : part + (size_t)s * k;   // single group: final value
    for (int32_t c = threadIdx.x * CPT; c < k; c += blockDim.x * CPT) {
        double acc[CPT];
#pragma unroll
        for (int u = 0; u < CPT; u++) acc[u] = 0.0;      // 0 + v == v exactly: the first addend is taken as is
        int32_t t = b;
        for (; t + 4 <= e; t += 4) {                    // 4 independent gathers in flight, adds stay in order
            double f[4][CPT], sc[4];
#pragma unroll
            for (int q = 0; q < 4; q++) {
                const double* __restrict__ src = F + (size_t)(uint32_t)keys[t + q] * k + c;
                if (CPT == 2) { const double2 v = *reinterpret_cast<const double2*>(src); f[q][0] = v.x; f[q][CPT - 1] = v.y; }
                else f[q][0] = *src;
                sc[q] = (double)score[t + q];
            }
#pragma unroll
            for (int q = 0; q < 4; q++)
#pragma unroll
                for (int u = 0; u < CPT; u++) acc[u] = __dadd_rn(acc[u], __dmul_rn(f[q][u], sc[q]));
        }
        for (; t < e; t++) {
            const double* __restrict__ src = F + (size_t)(uint32_t)keys[t] * k + c;
            const double sc = (double)score[t];
#pragma unroll
            for (int u = 0; u < CPT; u++) acc[u] = __dadd_rn(acc[u], __dmul_rn(src[u], sc));
        }
#pragma unroll
        for (int u = 0; u < CPT; u++) dst[c + u] = acc[u];
    }
}

__global__ void k_join_combine(const int32_t* __restrict__ seg_off, int32_t n_rows, int32_t k,
                               const double* __restrict__ part, double* __restrict__ X) {
    const int32_t r = blockIdx.x * blockDim.y + threadIdx.y;
    if (r >= n_rows) return;
    const int32_t s0 = seg_off[r], s1 = seg_off[r + 1];
    if (s1 - s0 <= 1) return;                   // written by k_join_partial
    for (int32_t c = threadIdx.x; c < k; c += blockDim.x) {
        double acc = part[(size_t)s0 * k + c];
        for (int32_t s = s0 + 1; s < s1; s++) acc = __dadd_rn(acc, part[(size_t)s * k + c]);
        X[(size_t)r * k + c] = acc;
    }
}

// ---------------------------------------------------------------------------------------------
// C = sum_r F[r]^T F[r]   (CrossProductMapper + MatrixSumReducer as combiner + reducer)
// grid (splits, entry groups); thread owns CROSS_EPT entries (a, b) of the k x k matrix.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(CROSS_THREADS)
k_cross_partial(const double* __restrict__ F, int32_t n_rows, int32_t k, int32_t split_rows, double* __restrict__ part) {
    extern __shared__ double s_rows[];           // [CROSS_TILE][k]
    const int32_t kk = k * k;
    const int32_t r0 = blockIdx.x * split_rows, r1 = min(r0 + split_rows, n_rows);
    int32_t ea[CROSS_EPT], eb[CROSS_EPT];
    double acc[CROSS_EPT];
    const int32_t q0 = blockIdx.y * (CROSS_THREADS * CROSS_EPT) + threadIdx.x;
#pragma unroll
    for (int e = 0; e < CROSS_EPT; e++) {
        const int32_t q = q0 + e * CROSS_THREADS;
        ea[e] = (q < kk) ? q / k : 0; eb[e] = (q < kk) ? q % k : 0; acc[e] = 0.0;
    }
    for (int32_t t0 = r0; t0 < r1; t0 += CROSS_TILE) {
        const int32_t nt = min(CROSS_TILE, r1 - t0);
        __syncthreads();
        for (int32_t x = threadIdx.x; x < nt * k; x += CROSS_THREADS) s_rows[x] = F[(size_t)t0 * k + x];
        __syncthreads();
        for (int32_t t = 0; t < nt; t++) {
            const double* __restrict__ f = s_rows + t * k;
#pragma unroll
            for (int e = 0; e < CROSS_EPT; e++) acc[e] = __dadd_rn(acc[e], __dmul_rn(f[ea[e]], f[eb[e]]));
        }
    }
#pragma unroll
    for (int e = 0; e < CROSS_EPT; e++) {
        const int32_t q = q0 + e * CROSS_THREADS;
        if (q < kk) part[(size_t)blockIdx.x * kk + q] = acc[e];
    }
}

__global__ void k_cross_combine(const double* __restrict__ part, int32_t n_split, int32_t kk, double* __restrict__ Cm) {
    const int32_t q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= kk) return;
    double acc = part[q];
    for (int32_t s = 1; s < n_split; s++) acc = __dadd_rn(acc, part[(size_t)s * kk + q]);
    Cm[q] = acc;
}

// ---------------------------------------------------------------------------------------------
// update: y = C f (CHMapper.java:33-41 / WCMapper.java:37-48), then
//   kind 0  f .* x ./ (y + eps)                               HComputationReducer.java:57-75   (NMF H)
//   kind 1  d = f.y, e = f.x, f .* (x + d) ./ (y + e + eps)    PPCHComputationReducer.java:62-86 (PPC H, inf guard)
//   kind 2  f .* x ./ (y + eps) with the inf guard            WComputationMapper.java:100-114   (W)
// blockDim = (TX, R): R rows per block, TX threads stride the k columns.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ double guard_inf(double v) { return isinf(v) ? DBL_MAX : v; }

__global__ void __launch_bounds__(256)
k_update(int kind, const double* __restrict__ F, const double* __restrict__ X, const double* __restrict__ Cm,
         int32_t n_rows, int32_t k, int do_norm, double* __restrict__ out) {
    extern __shared__ double s_upd[];            // [R][3][k]: f, x, y
    const int32_t r = blockIdx.x * blockDim.y + threadIdx.y;
    double* __restrict__ sf = s_upd + (size_t)threadIdx.y * 3 * k;
    double* __restrict__ sx = sf + k;
    double* __restrict__ sy = sx + k;
    const bool live = r < n_rows;
    if (live)
        for (int32_t c = threadIdx.x; c < k; c += blockDim.x) { sf[c] = F[(size_t)r * k + c]; sx[c] = X[(size_t)r * k + c]; }
    __syncthreads();
    if (live)
        for (int32_t c = threadIdx.x; c < k; c += blockDim.x) {
            double s = 0.0;
            for (int32_t q = 0; q < k; q++) s = __dadd_rn(s, __dmul_rn(Cm[(size_t)q * k + c], sf[q]));   // C is bit-symmetric
            sy[c] = s;
        }
    __syncthreads();
    double d = 0.0, e = 0.0;
    if (live && kind == 1) {
        for (int32_t q = 0; q < k; q++) d = __dadd_rn(d, __dmul_rn(sf[q], sy[q]));      // vectorH.dot(vectorY)
        for (int32_t q = 0; q < k; q++) e = __dadd_rn(e, __dmul_rn(sf[q], sx[q]));      // vectorH.dot(vectorX)
    }
    __syncthreads();                             // sy is overwritten below when normalising
    if (live)
        for (int32_t c = threadIdx.x; c < k; c += blockDim.x) {
            double a = sx[c], b = sy[c];
            if (kind == 1) { a = guard_inf(__dadd_rn(a, d)); b = guard_inf(__dadd_rn(b, e)); }
            else if (kind == 2) { a = guard_inf(a); b = guard_inf(b); }
            const double o = __dmul_rn(sf[c], __ddiv_rn(a, __dadd_rn(b, NMF_EPS)));
            if (do_norm) sx[c] = o; else out[(size_t)r * k + c] = o;
        }
    if (do_norm) {                               // intended L1 renormalisation (not what the reference does)
        __syncthreads();
        if (live) {
            double n1 = 0.0;
            for (int32_t q = 0; q < k; q++) n1 = __dadd_rn(n1, fabs(sx[q]));
            for (int32_t c = threadIdx.x; c < k; c += blockDim.x) out[(size_t)r * k + c] = __ddiv_rn(sx[c], n1);
        }
    }
}

// FindClusterMapper.java:34-42 (Mahout 0.8 maxValueIndex) + CountReducer.java:35-45
__global__ void k_argmax(const double* __restrict__ H, int32_t n_rows, int32_t k, int32_t* __restrict__ cluster,
                         int32_t* __restrict__ count) {
    const int32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_rows) return;
    int best = -1, nz = 0, first_zero = -1;
    double mx = -INFINITY;
    for (int32_t c = 0; c < k; c++) {
        const double v = H[(size_t)r * k + c];
        if (v == 0.0) { if (first_zero < 0) first_zero = c; continue; }
        nz++;
        if (v > mx) { mx = v; best = c; }
    }
    if (nz < k && mx < 0.0) best = first_zero;
    cluster[r] = best;
    if (best >= 0) atomicAdd(&count[best], 1);
}

// DataInitialization.java:79-88: nextDouble() + eps per element, row divided by its L1 norm
__device__ __forceinline__ uint64_t splitmix64(uint64_t x) {
    x += 0x9e3779b97f4a7c15ull;
    x = (x ^ (x >> 30)) * 0xbf58476d1ce4e5b9ull;
    x = (x ^ (x >> 27)) * 0x94d049bb133111ebull;
    return x ^ (x >> 31);
}
__global__ void k_random_rows(double* __restrict__ F, int32_t n_rows, int32_t k, uint64_t seed) {
    const int32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_rows) return;
    double n1 = 0.0;
    for (int32_t c = 0; c < k; c++) {
        const uint64_t bits = splitmix64(seed ^ splitmix64((uint64_t)r * (uint64_t)k + (uint64_t)c));
        const double v = (double)(bits >> 11) * (1.0 / 9007199254740992.0) + NMF_EPS;    // 53-bit uniform, as nextDouble()
        F[(size_t)r * k + c] = v;
        n1 = __dadd_rn(n1, fabs(v));
    }
    for (int32_t c = 0; c < k; c++) F[(size_t)r * k + c] = __ddiv_rn(F[(size_t)r * k + c], n1);
}

}  // namespace

struct fy_nmf_ctx {
    fy_nmf_params prm{};
    char err[512] = {0};
    cudaStream_t stream = nullptr;
    int64_t launches = 0;

    int64_t nnz = 0;
    NBuf<int32_t> in_user, in_item;
    NBuf<float> in_score;
    bool have_ratings = false, have_factors = false, indexed = false;

    // index: [0] by user (CSR), [1] by item (CSC)
    int32_t m = 0;
    NBuf<uint64_t> key_raw[2], key_sorted[2];
    NBuf<float> sc_sorted[2];
    NBuf<int32_t> rowptr[2], nseg[2], seg_off[2], seg_row[2];
    int32_t n_seg[2] = {0, 0};
    NBuf<unsigned char> cub_tmp;
    NBuf<unsigned long long> counters;
    NBuf<int> flags, first_empty;

    // factors (ping-pong) and per-iteration scratch
    NBuf<double> H[2], W[2], XH, XW, CW, CH, part_join[2], part_cross;
    int cur = 0;
    NBuf<int32_t> cl, cl_count;

    cudaGraphExec_t graph[2] = {nullptr, nullptr};
    fy_nmf_profile prof{};
    cudaEvent_t ev[3] = {nullptr, nullptr, nullptr};

    int fail(int code, const char* fmt, ...) {
        va_list ap; va_start(ap, fmt); vsnprintf(err, sizeof(err), fmt, ap); va_end(ap);
        return code;
    }
    void drop_graphs() {
        for (auto& g : graph) if (g) { cudaGraphExecDestroy(g); g = nullptr; }
    }
};

#define NLAUNCH(ctx, kernel, grid, block, smem, ...)                                  \
    do {                                                                              \
        kernel<<<(grid), (block), (smem), (ctx)->stream>>>(__VA_ARGS__);              \
        (ctx)->launches++;                                                            \
        NCK(cudaGetLastError());                                                      \
    } while (0)

template <class Fn>
static int nmf_guarded(fy_nmf_ctx* ctx, Fn&& f) {
    try {
        return f();
    } catch (const NmfCudaFail& c) {
        const int code = (c.err == cudaErrorMemoryAllocation) ? FY_E_NOMEM : FY_E_CUDA;
        return ctx->fail(code, "CUDA error %d (%s) at %s, nmf_engine.cu:%d", (int)c.err, cudaGetErrorString(c.err), c.what, c.line);
    } catch (const std::bad_alloc&) {
        return ctx->fail(FY_E_NOMEM, "host allocation failed");
    } catch (...) {
        return ctx->fail(FY_E_CUDA, "unexpected exception");
    }
}

extern "C" void fy_nmf_default_params(fy_nmf_params* p) {
    if (!p) return;
    std::memset(p, 0, sizeof(*p));
    p->mode = 1;                          // RMRecommenderDriver runs PPCDriver (RMRecommenderDriver.java:170-176)
    p->number_of_iterations = 10;         // :94
    p->normalization_frequency = 12;      // :115
    p->apply_normalization = 0;
    p->id_base = 1;
    p->combine_len = 1024;
    p->split_rows = 256;
    p->device = 0;
}

extern "C" int fy_nmf_create(fy_nmf_ctx** out, const fy_nmf_params* p) {
    if (!out || !p) return FY_E_ARG;
    *out = nullptr;
    if ((p->mode != 0 && p->mode != 1) || p->number_of_users <= 0 || p->number_of_items <= 0 || p->number_of_clusters <= 0 ||
        p->number_of_iterations < 0 || p->combine_len < 0 || p->split_rows < 0)
        return FY_E_ARG;
    if (p->number_of_clusters > MAX_K) return FY_E_UNSUPPORTED;
    int n_dev = 0;
    if (cudaGetDeviceCount(&n_dev) != cudaSuccess || n_dev <= 0 || p->device < 0 || p->device >= n_dev)
        return FY_E_CUDA;                 // no CPU fallback: fail loudly
    fy_nmf_ctx* ctx = new (std::nothrow) fy_nmf_ctx();
    if (!ctx) return FY_E_NOMEM;
    ctx->prm = *p;
    int rc = nmf_guarded(ctx, [&]() {
        NCK(cudaSetDevice(p->device));
        cudaDeviceProp prop;
        NCK(cudaGetDeviceProperties(&prop, p->device));
        if (prop.major < 10) return ctx->fail(FY_E_CUDA, "device %d is sm_%d%d; this library is built for sm_100a only", p->device, prop.major, prop.minor);
        NCK(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
        for (auto& e : ctx->ev) NCK(cudaEventCreate(&e));
        return (int)FY_OK;
    });
    if (rc != FY_OK) { fy_nmf_destroy(ctx); return rc; }
    *out = ctx;
    return FY_OK;
}

extern "C" void fy_nmf_destroy(fy_nmf_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->prm.device);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    ctx->drop_graphs();
    for (auto& e : ctx->ev) if (e) cudaEventDestroy(e);
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

extern "C" const char* fy_nmf_last_error(const fy_nmf_ctx* ctx) { return ctx ? ctx->err : "null context"; }

extern "C" int fy_nmf_set_ratings(fy_nmf_ctx* ctx, const int32_t* user, const int32_t* item, const float* score, int64_t nnz) {
    if (!ctx) return FY_E_ARG;
    if (!user || !item || !score || nnz <= 0) return ctx->fail(FY_E_ARG, "fy_nmf_set_ratings: null pointer or nnz <= 0");
    if (nnz > 0x7fffffffll - 1024) return ctx->fail(FY_E_UNSUPPORTED, "more than 2^31 ratings");
    return nmf_guarded(ctx, [&]() {
        NCK(cudaSetDevice(ctx->prm.device));
        ctx->have_ratings = false; ctx->indexed = false;
        ctx->in_user.need((size_t)nnz); ctx->in_item.need((size_t)nnz); ctx->in_score.need((size_t)nnz);
        NCK(cudaMemcpyAsync(ctx->in_user.p, user, (size_t)nnz * 4, cudaMemcpyHostToDevice, ctx->stream));
        NCK(cudaMemcpyAsync(ctx->in_item.p, item, (size_t)nnz * 4, cudaMemcpyHostToDevice, ctx->stream));
        NCK(cudaMemcpyAsync(ctx->in_score.p, score, (size_t)nnz * 4, cudaMemcpyHostToDevice, ctx->stream));
        NCK(cudaStreamSynchronize(ctx->stream));
        ctx->nnz = nnz;
        ctx->have_ratings = true;
        return (int)FY_OK;
    });
}

static void alloc_factors(fy_nmf_ctx* ctx) {
    const size_t k = (size_t)ctx->prm.number_of_clusters;
    for (int b = 0; b < 2; b++) { ctx->H[b].need((size_t)ctx->prm.number_of_users * k); ctx->W[b].need((size_t)ctx->prm.number_of_items * k); }
}

extern "C" int fy_nmf_set_factors(fy_nmf_ctx* ctx, const double* H, const double* W) {
    if (!ctx) return FY_E_ARG;
    if (!H || !W) return ctx->fail(FY_E_ARG, "fy_nmf_set_factors: null pointer");
    return nmf_guarded(ctx, [&]() {
        NCK(cudaSetDevice(ctx->prm.device));
        alloc_factors(ctx);
        const size_t k = (size_t)ctx->prm.number_of_clusters;
        ctx->cur = 0;
        NCK(cudaMemcpyAsync(ctx->H[0].p, H, (size_t)ctx->prm.number_of_users * k * 8, cudaMemcpyHostToDevice, ctx->stream));
        NCK(cudaMemcpyAsync(ctx->W[0].p, W, (size_t)ctx->prm.number_of_items * k * 8, cudaMemcpyHostToDevice, ctx->stream));
        NCK(cudaStreamSynchronize(ctx->stream));
        ctx->have_factors = true;
        return (int)FY_OK;
    });
}

extern "C" int fy_nmf_init_random(fy_nmf_ctx* ctx, uint64_t seed) {
    if (!ctx) return FY_E_ARG;
    return nmf_guarded(ctx, [&]() {
        NCK(cudaSetDevice(ctx->prm.device));
        alloc_factors(ctx);
        ctx->cur = 0;
        const int32_t k = ctx->prm.number_of_clusters;
        NLAUNCH(ctx, k_random_rows, ncdiv(ctx->prm.number_of_users, 128), 128, 0, ctx->H[0].p, ctx->prm.number_of_users, k, seed);
        NLAUNCH(ctx, k_random_rows, ncdiv(ctx->prm.number_of_items, 128), 128, 0, ctx->W[0].p, ctx->prm.number_of_items, k,
                seed ^ 0x5851f42d4c957f2dull);
        NCK(cudaStreamSynchronize(ctx->stream));
        ctx->have_factors = true;
        return (int)FY_OK;
    });
}

// CSR (by user) and CSC (by item) with their combiner segments; once per set of ratings
static int build_index(fy_nmf_ctx* ctx) {
    cudaStream_t st = ctx->stream;
    const int64_t nnz = ctx->nnz;
    const int32_t U = ctx->prm.number_of_users, M = ctx->prm.number_of_items;
    ctx->counters.need(1); ctx->flags.need(NF_COUNT); ctx->first_empty.need(2);
    NCK(cudaMemsetAsync(ctx->counters.p, 0, sizeof(unsigned long long), st));
    NCK(cudaMemsetAsync(ctx->flags.p, 0, sizeof(int) * NF_COUNT, st));
    NCK(cudaMemsetAsync(ctx->first_empty.p, 0x7f, sizeof(int) * 2, st));
    for (int o = 0; o < 2; o++) { ctx->key_raw[o].need((size_t)nnz); ctx->key_sorted[o].need((size_t)nnz); ctx->sc_sorted[o].need((size_t)nnz); }
    NLAUNCH(ctx, k_nmf_keys, ncdiv(nnz, 256), 256, 0, ctx->in_user.p, ctx->in_item.p, ctx->in_score.p, nnz, ctx->prm.id_base, U, M,
            ctx->key_raw[0].p, ctx->key_raw[1].p, ctx->counters.p, ctx->flags.p);
    for (int o = 0; o < 2; o++) {          // stable: equal (row, col) pairs keep their input order
        size_t tmp = 0;
        NCK(cub::DeviceRadixSort::SortPairs(nullptr, tmp, ctx->key_raw[o].p, ctx->key_sorted[o].p, ctx->in_score.p, ctx->sc_sorted[o].p, nnz, 0, 64, st));
        ctx->cub_tmp.need(tmp);
        NCK(cub::DeviceRadixSort::SortPairs(ctx->cub_tmp.p, tmp, ctx->key_raw[o].p, ctx->key_sorted[o].p, ctx->in_score.p, ctx->sc_sorted[o].p, nnz, 0, 64, st));
    }
    unsigned long long h_valid = 0; int h_flags[NF_COUNT];
    NCK(cudaMemcpyAsync(&h_valid, ctx->counters.p, sizeof(h_valid), cudaMemcpyDeviceToHost, st));
    NCK(cudaMemcpyAsync(h_flags, ctx->flags.p, sizeof(h_flags), cudaMemcpyDeviceToHost, st));
    NCK(cudaStreamSynchronize(st));
    if (h_flags[NF_BAD_ID]) return ctx->fail(FY_E_ARG, "user or item id outside [id_base, id_base + n)");
    const int32_t m = (int32_t)h_valid;
    ctx->m = m;
    const int32_t n_rows[2] = {U, M};
    for (int o = 0; o < 2; o++) {
        const int32_t n = n_rows[o];
        ctx->rowptr[o].need((size_t)n + 1); ctx->nseg[o].need((size_t)n + 1); ctx->seg_off[o].need((size_t)n + 1);
        NLAUNCH(ctx, k_nmf_rowptr, ncdiv(n + 1, 256), 256, 0, ctx->key_sorted[o].p, m, n, ctx->rowptr[o].p);
        NCK(cudaMemsetAsync(ctx->nseg[o].p, 0, ((size_t)n + 1) * 4, st));
        NLAUNCH(ctx, k_nmf_segcount, ncdiv(n, 256), 256, 0, ctx->rowptr[o].p, n, ctx->prm.combine_len, ctx->nseg[o].p, ctx->first_empty.p + o);
        size_t tmp = 0;
        NCK(cub::DeviceScan::ExclusiveSum(nullptr, tmp, ctx->nseg[o].p, ctx->seg_off[o].p, n + 1, st));
        ctx->cub_tmp.need(tmp);
        NCK(cub::DeviceScan::ExclusiveSum(ctx->cub_tmp.p, tmp, ctx->nseg[o].p, ctx->seg_off[o].p, n + 1, st));
    }
    int h_empty[2]; int32_t h_nseg[2];
    NCK(cudaMemcpyAsync(h_empty, ctx->first_empty.p, sizeof(h_empty), cudaMemcpyDeviceToHost, st));
    for (int o = 0; o < 2; o++) NCK(cudaMemcpyAsync(&h_nseg[o], ctx->seg_off[o].p + n_rows[o], 4, cudaMemcpyDeviceToHost, st));
    NCK(cudaStreamSynchronize(st));
    if (h_empty[0] < U) return ctx->fail(FY_E_USER_WITHOUT_RATING, "User %d has not rated any item", h_empty[0] + ctx->prm.id_base);       // HComputationReducer.java:52-55
    if (h_empty[1] < M) return ctx->fail(FY_E_ITEM_WITHOUT_RATING, "Item %d has not been rated by anybody", h_empty[1] + ctx->prm.id_base); // WComputationMapper.java:95-98
    const size_t k = (size_t)ctx->prm.number_of_clusters;
    for (int o = 0; o < 2; o++) {
        ctx->n_seg[o] = h_nseg[o];
        ctx->seg_row[o].need((size_t)h_nseg[o]);
        ctx->part_join[o].need((size_t)h_nseg[o] * k);
        NLAUNCH(ctx, k_nmf_segfill, ncdiv(n_rows[o], 256), 256, 0, ctx->seg_off[o].p, n_rows[o], ctx->seg_row[o].p);
    }
    ctx->indexed = true;
    return FY_OK;
}

// one iteration: (H[src], W[src]) -> (H[dst], W[dst]); kernels only, so it can be stream-captured
static void enqueue_iteration(fy_nmf_ctx* ctx, int src, int dst, int do_norm) {
    const int32_t U = ctx->prm.number_of_users, M = ctx->prm.number_of_items, k = ctx->prm.number_of_clusters;
    const int32_t kk = k * k;
    const int TX = std::min(((k + 31) / 32) * 32, 256), R = std::max(1, 256 / TX);
    const dim3 blk(TX, R);
    const int32_t split = ctx->prm.split_rows;
    const double* Hs = ctx->H[src].p; const double* Ws = ctx->W[src].p;
    struct Side { int o; const double* gather; double* X; int32_t n_rows; const double* cross_src; int32_t cross_rows; double* Cm;
                  const double* F; double* out; int kind; };
    // H side gathers W rows and needs C_W; W side gathers H rows and needs C_H
    const Side sides[2] = {
        {0, Ws, ctx->XH.p, U, Ws, M, ctx->CW.p, Hs, ctx->H[dst].p, ctx->prm.mode == 1 ? 1 : 0},
        {1, Hs, ctx->XW.p, M, Hs, U, ctx->CH.p, Ws, ctx->W[dst].p, 2},
    };
    // join: two factor columns per thread (one 16-byte gather per rating) when k is even and at least a warp's worth
    // of column pairs exists (measured: 0.84 -> 0.71 ms per side at k = 50; no gain at k = 10)
    const bool pairs = (k % 2 == 0) && k >= 32;
    const int jcols = pairs ? k / 2 : k;
    const int JX = std::min(((jcols + 31) / 32) * 32, 256), JR = std::max(1, 256 / JX);
    const dim3 jblk(JX, JR);
    for (const Side& s : sides) {
        if (pairs)
            NLAUNCH(ctx, k_join_partial<2>, ncdiv(ctx->n_seg[s.o], JR), jblk, 0, ctx->seg_row[s.o].p, ctx->seg_off[s.o].p, ctx->n_seg[s.o],
                    ctx->rowptr[s.o].p, ctx->key_sorted[s.o].p, ctx->sc_sorted[s.o].p, s.gather, k, ctx->prm.combine_len,
                    ctx->part_join[s.o].p, s.X);
        else
            NLAUNCH(ctx, k_join_partial<1>, ncdiv(ctx->n_seg[s.o], JR), jblk, 0, ctx->seg_row[s.o].p, ctx->seg_off[s.o].p, ctx->n_seg[s.o],
                    ctx->rowptr[s.o].p, ctx->key_sorted[s.o].p, ctx->sc_sorted[s.o].p, s.gather, k, ctx->prm.combine_len,
                    ctx->part_join[s.o].p, s.X);
        if (ctx->n_seg[s.o] > s.n_rows)
            NLAUNCH(ctx, k_join_combine, ncdiv(s.n_rows, R), blk, 0, ctx->seg_off[s.o].p, s.n_rows, k, ctx->part_join[s.o].p, s.X);
        const int32_t sr = split > 0 ? split : s.cross_rows;
        const int32_t n_split = ncdiv(s.cross_rows, sr);
        NLAUNCH(ctx, k_cross_partial, dim3(n_split, ncdiv(kk, CROSS_THREADS * CROSS_EPT)), CROSS_THREADS, (size_t)CROSS_TILE * k * 8,
                s.cross_src, s.cross_rows, k, sr, ctx->part_cross.p);
        NLAUNCH(ctx, k_cross_combine, ncdiv(kk, 256), 256, 0, ctx->part_cross.p, n_split, kk, s.Cm);
        NLAUNCH(ctx, k_update, ncdiv(s.n_rows, R), blk, (size_t)R * 3 * k * 8, s.kind, s.F, s.X, s.Cm, s.n_rows, k,
                (s.kind == 1) ? do_norm : 0, s.out);
    }
}

extern "C" int fy_nmf_run(fy_nmf_ctx* ctx) {
    if (!ctx) return FY_E_ARG;
    if (!ctx->have_ratings || !ctx->have_factors) return ctx->fail(FY_E_STATE, "fy_nmf_run needs fy_nmf_set_ratings and fy_nmf_set_factors / fy_nmf_init_random first");
    return nmf_guarded(ctx, [&]() {
        NCK(cudaSetDevice(ctx->prm.device));
        cudaStream_t st = ctx->stream;
        ctx->launches = 0;
        ctx->prof = fy_nmf_profile{};
        const int32_t U = ctx->prm.number_of_users, M = ctx->prm.number_of_items, k = ctx->prm.number_of_clusters;
        NCK(cudaEventRecord(ctx->ev[0], st));
        if (!ctx->indexed) { ctx->drop_graphs(); const int rc = build_index(ctx); if (rc != FY_OK) return rc; }
        ctx->XH.need((size_t)U * k); ctx->XW.need((size_t)M * k); ctx->CW.need((size_t)k * k); ctx->CH.need((size_t)k * k);
        const int32_t sr = ctx->prm.split_rows;
        const int32_t max_split = std::max(sr > 0 ? ncdiv(U, sr) : 1, sr > 0 ? ncdiv(M, sr) : 1);
        const size_t old_cap = ctx->part_cross.cap;
        ctx->part_cross.need((size_t)max_split * k * k);
        if (ctx->part_cross.cap != old_cap) ctx->drop_graphs();
        if ((size_t)CROSS_TILE * k * 8 > 48 * 1024) NCK(cudaFuncSetAttribute(k_cross_partial, cudaFuncAttributeMaxDynamicSharedMemorySize, CROSS_TILE * k * 8));
        NCK(cudaEventRecord(ctx->ev[1], st));

        const int n_iter = ctx->prm.number_of_iterations;
        const int nf = ctx->prm.normalization_frequency;
        // `iteration % normalizationFrequency == 0` (PPCHComputationReducer.java:88); constant for +-1
        const bool norm_on = ctx->prm.apply_normalization != 0 && ctx->prm.mode == 1 && nf != 0;
        const bool norm_const = !norm_on || nf == 1 || nf == -1;
        const bool use_graph = norm_const && n_iter >= 4;
        if (use_graph && !(ctx->graph[0] && ctx->graph[1])) {
            // the launch-bound inner loop (8-10 small kernels) is captured once per direction of the ping-pong; a failure
            // half way (direction 0 captured, direction 1 not) must not leave a null handle behind for the next run
            ctx->drop_graphs();
            try {
                for (int dir = 0; dir < 2; dir++) {
                    cudaGraph_t g = nullptr;
                    const int64_t keep = ctx->launches;
                    NCK(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
                    try { enqueue_iteration(ctx, dir, dir ^ 1, norm_on ? 1 : 0); }
                    catch (...) { cudaStreamEndCapture(st, &g); if (g) cudaGraphDestroy(g); throw; }
                    NCK(cudaStreamEndCapture(st, &g));
                    ctx->launches = keep;                 // captured, not launched
                    const cudaError_t e = cudaGraphInstantiate(&ctx->graph[dir], g, 0);
                    cudaGraphDestroy(g);
                    NCK(e);
                }
            } catch (...) { ctx->drop_graphs(); throw; }
        }
        int64_t graph_kernels = 0;
        for (int it = 1; it <= n_iter; it++) {
            const int src = ctx->cur, dst = ctx->cur ^ 1;
            if (use_graph) {
                NCK(cudaGraphLaunch(ctx->graph[src], st));
                ctx->prof.graph_replays++;
                graph_kernels++;
            } else {
                enqueue_iteration(ctx, src, dst, (norm_on && it % nf == 0) ? 1 : 0);
            }
            ctx->cur = dst;
        }
        if (graph_kernels) {
            // kernels per captured iteration: join partial (+combine when a row has several groups), cross partial + combine, update; x2
            int64_t per = 0;
            for (int o = 0; o < 2; o++) per += 4 + ((ctx->n_seg[o] > (o == 0 ? U : M)) ? 1 : 0);
            ctx->launches += per * graph_kernels;
        }
        NCK(cudaEventRecord(ctx->ev[2], st));
        NCK(cudaStreamSynchronize(st));
        float ms = 0;
        NCK(cudaEventElapsedTime(&ms, ctx->ev[0], ctx->ev[1])); ctx->prof.ms_index = ms;
        NCK(cudaEventElapsedTime(&ms, ctx->ev[1], ctx->ev[2])); ctx->prof.ms_iterations = ms;
        ctx->prof.ms_per_iteration = n_iter > 0 ? ms / n_iter : 0.0;
        ctx->prof.iterations = n_iter;
        ctx->prof.ratings = ctx->m;
        ctx->prof.join_bytes = 2.0 * (double)ctx->m * (8.0 * k + 12.0);
        ctx->prof.kernel_launches = ctx->launches;
        return (int)FY_OK;
    });
}

extern "C" int fy_nmf_get_factors(fy_nmf_ctx* ctx, double* H, double* W) {
    if (!ctx) return FY_E_ARG;
    if (!ctx->have_factors) return ctx->fail(FY_E_STATE, "no factors yet");
    return nmf_guarded(ctx, [&]() {
        NCK(cudaSetDevice(ctx->prm.device));
        const size_t k = (size_t)ctx->prm.number_of_clusters;
        if (H) NCK(cudaMemcpyAsync(H, ctx->H[ctx->cur].p, (size_t)ctx->prm.number_of_users * k * 8, cudaMemcpyDeviceToHost, ctx->stream));
        if (W) NCK(cudaMemcpyAsync(W, ctx->W[ctx->cur].p, (size_t)ctx->prm.number_of_items * k * 8, cudaMemcpyDeviceToHost, ctx->stream));
        NCK(cudaStreamSynchronize(ctx->stream));
        return (int)FY_OK;
    });
}

extern "C" int fy_nmf_cluster_assignment(fy_nmf_ctx* ctx, int32_t* cluster_out, int32_t* cluster_size_out) {
    if (!ctx) return FY_E_ARG;
    if (!ctx->have_factors) return ctx->fail(FY_E_STATE, "no factors yet");
    if (!cluster_out) return ctx->fail(FY_E_ARG, "fy_nmf_cluster_assignment: null pointer");
    return nmf_guarded(ctx, [&]() {
        NCK(cudaSetDevice(ctx->prm.device));
        const int32_t U = ctx->prm.number_of_users, k = ctx->prm.number_of_clusters;
        ctx->cl.need(U); ctx->cl_count.need(k);
        NCK(cudaMemsetAsync(ctx->cl_count.p, 0, (size_t)k * 4, ctx->stream));
        NLAUNCH(ctx, k_argmax, ncdiv(U, 128), 128, 0, ctx->H[ctx->cur].p, U, k, ctx->cl.p, ctx->cl_count.p);
        NCK(cudaMemcpyAsync(cluster_out, ctx->cl.p, (size_t)U * 4, cudaMemcpyDeviceToHost, ctx->stream));
        if (cluster_size_out) NCK(cudaMemcpyAsync(cluster_size_out, ctx->cl_count.p, (size_t)k * 4, cudaMemcpyDeviceToHost, ctx->stream));
        NCK(cudaStreamSynchronize(ctx->stream));
        return (int)FY_OK;
    });
}

extern "C" int fy_nmf_get_profile(const fy_nmf_ctx* ctx, fy_nmf_profile* out) {
    if (!ctx || !out) return FY_E_ARG;
    *out = ctx->prof;
    return FY_OK;
}
