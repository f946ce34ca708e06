// cooc_tcgen05.cu -- config 3: item-item co-occurrence counts C = B^T B on the
// binarised rating matrix, as an int8 tcgen05 GEMM with int32 accumulators in TMEM (bit-exact
// integer counts).  Replaces the arithmetic of Mahout's RowSimilarityJob(CooccurrenceCountSimilarity)
// called at M/baselinerecommender/BaselineRecommenderJob.java:241-253.
//
// Operands: Bt [n_items x K_pad] uint8 (K = user ids, contiguous, zero padded to 128) is both the
// A operand (128-row tiles) and the B operand (256-row tiles), K-major, staged by TMA with the
// 128-byte swizzle.  One CTA computes one 128 x 256 tile of C:
//   warp 0 : TMA producer (one elected lane), 4-stage ring of {A 16 KB, B 32 KB} with full/empty
//            mbarriers
//   warp 1 : MMA issuer (one elected lane): 4 x tcgen05.mma.cta_group::1.kind::i8 (M128 N256 K32)
//            per stage, tcgen05.commit frees the stage / publishes the accumulator
//   warp 2 : TMEM allocator (256 columns = 128 lanes x 256 x int32)
//   warps 4-7 : epilogue, tcgen05.ld 32x32b.x32 -> registers -> 128-byte row segments of C and of C^T
// Only tiles touching the upper triangle are computed (C is symmetric); the grid is rasterised in
// groups of 16 row-tiles so that operand panels are shared through L2.
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include <vector>

namespace cooc {

constexpr int BM = 128, BN = 256, BK = 128;          // BK in bytes = int8 elements = one swizzle row
constexpr int STAGES = 4;
constexpr int GROUP_M = 16;                          // row-tiles per rasterisation group
constexpr int UMMA_K = 32;                           // K per tcgen05.mma for 8-bit operands
constexpr uint32_t A_BYTES = BM * BK, B_BYTES = BN * BK, STAGE_BYTES = A_BYTES + B_BYTES;
constexpr int NUM_THREADS = 256;
constexpr int TMEM_COLS = 256;
constexpr size_t SMEM_BYTES = (size_t)STAGES * STAGE_BYTES + 1024 /*alignment slack*/ + 256 /*barriers*/;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile("{\n\t.reg .pred p;\n\t"
                     "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                     "selp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    } while (!done);
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* tm, uint32_t bar, int32_t c0, int32_t c1) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(dst), "l"((uint64_t)tm), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
// K-major operand tile with the 128-byte swizzle: rows 128 B apart, 8-row groups 1024 B apart
// (bit layout: cute::UMMA::SmemDescriptor, mma_sm100_desc.hpp)
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);          // start address       bits [ 0,14)
    d |= (uint64_t)1 << 16;                            // leading byte offset bits [16,30) (ignored for swizzled K-major)
    d |= (uint64_t)(1024u >> 4) << 32;                 // stride byte offset  bits [32,46)
    d |= (uint64_t)1 << 46;                            // descriptor version  bits [46,48) = 1 on sm_100
    d |= (uint64_t)2 << 61;                            // layout type         bits [61,64) = SWIZZLE_128B
    return d;
}
// instruction descriptor (cute::UMMA::InstrDescriptor): S32 accumulate, u8 x u8, K-major A and B
__device__ __forceinline__ constexpr uint32_t make_idesc_i8(int M, int N) {
    return (2u << 4) /*c_format = S32*/ | (0u << 7) /*a = u8*/ | (0u << 10) /*b = u8*/ |
           ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void mma_i8(uint32_t tmem_c, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\t"
                 "setp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t}"
                 ::"r"(tmem_c), "l"(da), "l"(db), "r"(idesc), "r"(accumulate), "r"(0u) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

__global__ void __launch_bounds__(NUM_THREADS, 1)
k_cooc_gemm(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
            int32_t* __restrict__ C, int m_rows, int n_items, int ldc, int num_k_blocks, int n_tiles_m, int n_tiles_n,
            int symmetric) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;                 // SWIZZLE_128B tiles need 1024-byte alignment
    const uint32_t bars = base + STAGES * STAGE_BYTES;
    auto sA = [&](int s) { return base + (uint32_t)s * STAGE_BYTES; };
    auto sB = [&](int s) { return base + (uint32_t)s * STAGE_BYTES + A_BYTES; };
    auto full = [&](int s) { return bars + 8u * (uint32_t)s; };
    auto empty = [&](int s) { return bars + 8u * (uint32_t)(STAGES + s); };
    const uint32_t tmem_full = bars + 8u * (2 * STAGES);
    const uint32_t tmem_slot = bars + 8u * (2 * STAGES + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // Rasterisation: consecutive CTAs walk GROUP_M row-tiles x all column-tiles, so the ~148 tiles in
    // flight share ~16 A row-tiles and ~9 B row-tiles and the operand panels stream from HBM about
    // once per wave instead of once per tile (Bt is GBs, L2 is 126 MB).
    const int tiles_per_super = GROUP_M * n_tiles_n;
    const int sid = (int)blockIdx.x / tiles_per_super, rem = (int)blockIdx.x % tiles_per_super;
    const int gm = min(GROUP_M, n_tiles_m - sid * GROUP_M);
    const int m_tile = sid * GROUP_M + rem % gm, n_tile = rem / gm;
    // C is symmetric: tiles entirely below the diagonal are produced by the mirrored store of their
    // transposes, so they are skipped (uniform exit before any barrier / TMEM allocation).
    if (symmetric && (n_tile + 1) * BN <= m_tile * BM) return;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&tmA) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&tmB) : "memory");
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < STAGES; s++) { mbar_init(full(s), 1); mbar_init(empty(s), 1); }
        mbar_init(tmem_full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"((uint32_t)TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    uint32_t tmem_base;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot) : "memory");

    if (warp == 0) {
        if (lane == 0) {                                           // ===== TMA producer =====
            for (int kb = 0; kb < num_k_blocks; kb++) {
                const int s = kb % STAGES;
                const uint32_t ph = (uint32_t)(kb / STAGES) & 1u;
                mbar_wait(empty(s), ph ^ 1u);
                mbar_expect_tx(full(s), STAGE_BYTES);
                tma_load_2d(sA(s), &tmA, full(s), kb * BK, m_tile * BM);
                tma_load_2d(sB(s), &tmB, full(s), kb * BK, n_tile * BN);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {                                           // ===== MMA issuer =====
            const uint32_t idesc = make_idesc_i8(BM, BN);
            for (int kb = 0; kb < num_k_blocks; kb++) {
                const int s = kb % STAGES;
                const uint32_t ph = (uint32_t)(kb / STAGES) & 1u;
                mbar_wait(full(s), ph);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
                for (int k = 0; k < BK / UMMA_K; k++) {
                    const uint64_t da = make_smem_desc(sA(s) + (uint32_t)k * UMMA_K);
                    const uint64_t db = make_smem_desc(sB(s) + (uint32_t)k * UMMA_K);
                    mma_i8(tmem_base, da, db, idesc, (kb | k) != 0 ? 1u : 0u);
                }
                umma_commit(empty(s));                             // stage free once these MMAs have read it
            }
            umma_commit(tmem_full);                                // accumulator complete
        }
    } else if (warp >= 4) {                                        // ===== epilogue =====
        mbar_wait(tmem_full, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const int q = warp & 3;                                    // TMEM lane quarter of this warp
        const int row = q * 32 + lane;
        const int m = m_tile * BM + row;
        int32_t* __restrict__ crow = C + (size_t)m * ldc + (size_t)n_tile * BN;
#pragma unroll 1
        for (int c = 0; c < BN / 32; c++) {
            uint32_t v[32];
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(c * 32);
            asm volatile(
                "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                  "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
                  "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
                  "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                : "r"(taddr) : "memory");
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            if (m < m_rows) {
#pragma unroll
                for (int x = 0; x < 8; x++) {
                    int4 o = make_int4((int)v[4 * x], (int)v[4 * x + 1], (int)v[4 * x + 2], (int)v[4 * x + 3]);
                    *reinterpret_cast<int4*>(crow + c * 32 + 4 * x) = o;
                }
            }
            // mirrored store C[n][m] = C[m][n]: lanes hold consecutive m, so each store is one
            // contiguous 128-byte segment of row n
            if (symmetric) {
                const int nb = n_tile * BN + c * 32;
#pragma unroll
                for (int x = 0; x < 32; x++)
                    if (nb + x < n_items) C[(size_t)(nb + x) * ldc + m] = (int32_t)v[x];
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 2) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)TMEM_COLS) : "memory");
    }
}

// ---------------------------------------------------------------------------------------------
// Second generation (round 2): PERSISTENT CTAs (one per SM) walking the rasterised tile sequence with a stride of
// gridDim.x, and TWO TMEM accumulators (2 x 256 columns = the whole 512-column TMEM of the SM): the epilogue of tile t
// (tcgen05.ld + 256 KB of stores, direct and mirrored) overlaps the MMAs of tile t + 1, the TMA ring keeps running
// across tile boundaries (no per-tile pipeline fill), and barrier init / TMEM allocation happen once per SM instead of
// once per tile.  Same operands, same MMA shape, same rasterisation and the same integer results as k_cooc_gemm.
//   warp 0 : TMA producer        warp 1 : MMA issuer        warp 2 : TMEM allocator        warps 4-7 : epilogue
//   barriers: full/empty per smem stage, tmem_full/tmem_empty per accumulator (tmem_empty counts the 4 epilogue warps)
// ---------------------------------------------------------------------------------------------
constexpr int TMEM_COLS2 = 512;

__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}

__global__ void __launch_bounds__(NUM_THREADS, 1)
k_cooc_gemm2(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
             int32_t* __restrict__ C, int m_rows, int n_items, int ldc, int num_k_blocks,
             const int2* __restrict__ tiles /* (m_tile, n_tile) of every tile to compute, rasterised order */, int total,
             int symmetric, int* __restrict__ next_tile /* zeroed before the launch */) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;                 // SWIZZLE_128B tiles need 1024-byte alignment
    const uint32_t bars = base + STAGES * STAGE_BYTES;
    auto sA = [&](int s) { return base + (uint32_t)s * STAGE_BYTES; };
    auto sB = [&](int s) { return base + (uint32_t)s * STAGE_BYTES + A_BYTES; };
    auto full = [&](int s) { return bars + 8u * (uint32_t)s; };
    auto empty = [&](int s) { return bars + 8u * (uint32_t)(STAGES + s); };
    auto tmem_full = [&](int a) { return bars + 8u * (uint32_t)(2 * STAGES + a); };
    auto tmem_empty = [&](int a) { return bars + 8u * (uint32_t)(2 * STAGES + 2 + a); };
    auto sched_full = [&](int a) { return bars + 8u * (uint32_t)(2 * STAGES + 4 + a); };
    auto sched_empty = [&](int a) { return bars + 8u * (uint32_t)(2 * STAGES + 6 + a); };
    const uint32_t tmem_slot = bars + 8u * (2 * STAGES + 8);
    const uint32_t sched_tile = bars + 8u * (2 * STAGES + 9);      // two int32 slots: the tile each accumulator works on (-1 = stop)

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // The tile list holds only the tiles that are computed (for a symmetric product: those touching the upper triangle), in
    // the rasterised order.  Tiles are handed out DYNAMICALLY: the producer takes the next list entry with one atomicAdd and
    // publishes it to the MMA issuer and the epilogue warps through a two-slot shared-memory mailbox (slot = accumulator).
    // The tiles in flight therefore stay a window of consecutive entries and operand panels are shared through L2 as with
    // the hardware scheduler of one-CTA-per-tile; a static stride let the CTAs drift apart (32-42 ms vs 28 at ML-20M shape).
    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&tmA) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&tmB) : "memory");
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < STAGES; s++) { mbar_init(full(s), 1); mbar_init(empty(s), 1); }
        for (int a = 0; a < 2; a++) { mbar_init(tmem_full(a), 1); mbar_init(tmem_empty(a), 4); mbar_init(sched_full(a), 1); mbar_init(sched_empty(a), 5); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"((uint32_t)TMEM_COLS2) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    uint32_t tmem_base;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot) : "memory");

    if (warp == 0) {
        if (lane == 0) {                                           // ===== TMA producer: the ring runs across tiles =====
            uint32_t it = 0;
            for (uint32_t tl = 0;; tl++) {
                const int slot = (int)(tl & 1u);
                mbar_wait(sched_empty(slot), ((tl >> 1) & 1u) ^ 1u);   // both consumers have read the previous entry of this slot
                int lin = atomicAdd(next_tile, 1);
                if (lin >= total) lin = -1;
                asm volatile("st.shared.s32 [%0], %1;" ::"r"(sched_tile + 4u * (uint32_t)slot), "r"(lin) : "memory");
                mbar_arrive(sched_full(slot));                         // release: the entry is visible to whoever acquires the barrier
                if (lin < 0) break;
                const int2 t = __ldg(tiles + lin);
                const int m_tile = t.x, n_tile = t.y;
                for (int kb = 0; kb < num_k_blocks; kb++, it++) {
                    const int s = (int)(it % STAGES);
                    const uint32_t ph = (it / STAGES) & 1u;
                    mbar_wait(empty(s), ph ^ 1u);
                    mbar_expect_tx(full(s), STAGE_BYTES);
                    tma_load_2d(sA(s), &tmA, full(s), kb * BK, m_tile * BM);
                    tma_load_2d(sB(s), &tmB, full(s), kb * BK, n_tile * BN);
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {                                           // ===== MMA issuer =====
            const uint32_t idesc = make_idesc_i8(BM, BN);
            uint32_t it = 0;
            for (uint32_t tl = 0;; tl++) {
                const int acc = (int)(tl & 1u);
                const uint32_t aph = (tl >> 1) & 1u;
                mbar_wait(sched_full(acc), aph);
                int lin;
                asm volatile("ld.shared.s32 %0, [%1];" : "=r"(lin) : "r"(sched_tile + 4u * (uint32_t)acc) : "memory");
                mbar_arrive(sched_empty(acc));
                if (lin < 0) break;
                mbar_wait(tmem_empty(acc), aph ^ 1u);              // the epilogue has drained this accumulator (first use: free)
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t tacc = tmem_base + (uint32_t)(acc * BN);
                for (int kb = 0; kb < num_k_blocks; kb++, it++) {
                    const int s = (int)(it % STAGES);
                    const uint32_t ph = (it / STAGES) & 1u;
                    mbar_wait(full(s), ph);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
                    for (int k = 0; k < BK / UMMA_K; k++) {
                        const uint64_t da = make_smem_desc(sA(s) + (uint32_t)k * UMMA_K);
                        const uint64_t db = make_smem_desc(sB(s) + (uint32_t)k * UMMA_K);
                        mma_i8(tacc, da, db, idesc, (kb | k) != 0 ? 1u : 0u);
                    }
                    umma_commit(empty(s));                         // stage free once these MMAs have read it
                }
                umma_commit(tmem_full(acc));                       // accumulator complete
            }
        }
    } else if (warp >= 4) {                                        // ===== epilogue =====
        const int q = warp & 3;                                    // TMEM lane quarter of this warp
        const int row = q * 32 + lane;
        for (uint32_t tl = 0;; tl++) {
            const int acc = (int)(tl & 1u);
            const uint32_t aph = (tl >> 1) & 1u;
            mbar_wait(sched_full(acc), aph);
            int lin;
            asm volatile("ld.shared.s32 %0, [%1];" : "=r"(lin) : "r"(sched_tile + 4u * (uint32_t)acc) : "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(sched_empty(acc));
            if (lin < 0) break;
            const int2 t = __ldg(tiles + lin);
            const int m_tile = t.x, n_tile = t.y;
            mbar_wait(tmem_full(acc), aph);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const int m = m_tile * BM + row;
            int32_t* __restrict__ crow = C + (size_t)m * ldc + (size_t)n_tile * BN;
#pragma unroll 1
            for (int c = 0; c < BN / 32; c++) {
                uint32_t v[32];
                const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BN + c * 32);
                asm volatile(
                    "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                    "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                    "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                    : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                      "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
                      "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
                      "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                    : "r"(taddr) : "memory");
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                if (m < m_rows) {
#pragma unroll
                    for (int x = 0; x < 8; x++) {
                        int4 o = make_int4((int)v[4 * x], (int)v[4 * x + 1], (int)v[4 * x + 2], (int)v[4 * x + 3]);
                        *reinterpret_cast<int4*>(crow + c * 32 + 4 * x) = o;
                    }
                }
                if (symmetric) {                                   // mirrored store: one 128-byte segment of row n per instruction
                    const int nb = n_tile * BN + c * 32;
#pragma unroll
                    for (int x = 0; x < 32; x++)
                        if (nb + x < n_items) C[(size_t)(nb + x) * ldc + m] = (int32_t)v[x];
                }
            }
            // this warp's quarter of the accumulator is in registers / stored: hand it back to the MMA issuer
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(tmem_empty(acc));
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 2) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)TMEM_COLS2) : "memory");
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

}  // namespace cooc

// Internal launchers (rm2_engine.cu).  C[a_rows x ldc] = A[a_rows x k_pad] * B[b_rows x k_pad]^T on uint8
// operands, int32 result; k_pad a multiple of 128, ldc a multiple of 256 and >= b_rows.  symmetric = 1
// (A == B): only tiles touching the upper triangle are computed and every tile is also stored mirrored.
extern "C" int fyi_gemm_u8_nt(const uint8_t* A, int a_rows, const uint8_t* B, int b_rows, int k_pad, int32_t* C, int ldc,
                              int symmetric, void* stream, char* err, size_t errlen) {
    using namespace cooc;
    static EncodeTiledFn encode = nullptr;
    if (!encode) {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
        if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || !fn) {
            snprintf(err, errlen, "cuTensorMapEncodeTiled not available (%d)", (int)e);
            return -7;
        }
        encode = (EncodeTiledFn)fn;
    }
    if (k_pad % BK != 0 || ldc % BN != 0 || ldc < b_rows || a_rows <= 0 || b_rows <= 0 || (symmetric && (A != B || a_rows != b_rows))) {
        snprintf(err, errlen, "bad GEMM geometry");
        return -1;
    }
    CUtensorMap tmA, tmB;
    const cuuint64_t gdimA[2] = {(cuuint64_t)k_pad, (cuuint64_t)a_rows}, gdimB[2] = {(cuuint64_t)k_pad, (cuuint64_t)b_rows};
    const cuuint64_t gstride[1] = {(cuuint64_t)k_pad};
    const cuuint32_t estr[2] = {1, 1};
    const cuuint32_t boxA[2] = {(cuuint32_t)BK, (cuuint32_t)BM}, boxB[2] = {(cuuint32_t)BK, (cuuint32_t)BN};
    CUresult r1 = encode(&tmA, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, (void*)A, gdimA, gstride, boxA, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    CUresult r2 = encode(&tmB, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, (void*)B, gdimB, gstride, boxB, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r1 != CUDA_SUCCESS || r2 != CUDA_SUCCESS) { snprintf(err, errlen, "cuTensorMapEncodeTiled failed (%d, %d)", (int)r1, (int)r2); return -7; }
    const int nt_n = (b_rows + BN - 1) / BN, nt_m = (a_rows + BM - 1) / BM;
    // Which kernel: measured on one B200 (tools/cooc_bench.py, profiles/r02_cooc_bench.json) the persistent kernel wins where a
    // tile is short (ML-1M shape, 48 k-blocks per tile: 55.3 vs 59.5 us) and loses where a tile is 1 083 k-blocks long (ML-20M
    // shape: 30.4-32 vs 28.2 ms -- the one-tile-per-CTA kernel already keeps the tensor pipe 79 % busy at the power cap, and
    // the overlapped epilogue only adds traffic).  FY_COOC_V1=1 / 0 forces one or the other.
    const char* v1 = getenv("FY_COOC_V1");
    const bool use_v1 = v1 ? (v1[0] == '1') : (k_pad / BK > 256);
    cudaError_t e;
    if (use_v1) {
        e = cudaFuncSetAttribute(k_cooc_gemm, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES);
        if (e != cudaSuccess) { snprintf(err, errlen, "cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return -7; }
        k_cooc_gemm<<<nt_n * nt_m, NUM_THREADS, SMEM_BYTES, (cudaStream_t)stream>>>(tmA, tmB, C, a_rows, b_rows, ldc, k_pad / BK, nt_m, nt_n,
                                                                                  symmetric);
    } else {
        static int n_sm = 0;
        if (!n_sm) { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev); if (n_sm <= 0) n_sm = 148; }
        // tile list of this geometry (host-built, cached per device; 8 bytes per tile)
        struct TileCache { int dev = -1, nt_m = 0, nt_n = 0, sym = -1, count = 0; int2* d = nullptr; int* counter = nullptr; };
        static TileCache cache[16];
        int dev = 0;
        cudaGetDevice(&dev);
        TileCache& tc = cache[dev & 15];
        if (tc.dev != dev || tc.nt_m != nt_m || tc.nt_n != nt_n || tc.sym != symmetric) {
            std::vector<int2> h;
            const int tiles_per_super = GROUP_M * nt_n;
            for (int lin = 0; lin < nt_m * nt_n; lin++) {
                const int sid = lin / tiles_per_super, rem = lin % tiles_per_super;
                const int gm = (GROUP_M < nt_m - sid * GROUP_M) ? GROUP_M : nt_m - sid * GROUP_M;
                const int m_tile = sid * GROUP_M + rem % gm, n_tile = rem / gm;
                if (symmetric && (n_tile + 1) * BN <= m_tile * BM) continue;      // entirely below the diagonal: mirrored store of its transpose
                h.push_back(make_int2(m_tile, n_tile));
            }
            if (tc.d) { cudaStreamSynchronize((cudaStream_t)stream); cudaFree(tc.d); tc.d = nullptr; }
            e = cudaMalloc((void**)&tc.d, h.size() * sizeof(int2));
            if (e != cudaSuccess) { snprintf(err, errlen, "cudaMalloc(tile list): %s", cudaGetErrorString(e)); return -6; }
            e = cudaMemcpy(tc.d, h.data(), h.size() * sizeof(int2), cudaMemcpyHostToDevice);
            if (e != cudaSuccess) { snprintf(err, errlen, "cudaMemcpy(tile list): %s", cudaGetErrorString(e)); return -7; }
            tc.dev = dev; tc.nt_m = nt_m; tc.nt_n = nt_n; tc.sym = symmetric; tc.count = (int)h.size();
        }
        if (!tc.counter) {
            e = cudaMalloc((void**)&tc.counter, sizeof(int));
            if (e != cudaSuccess) { snprintf(err, errlen, "cudaMalloc(tile counter): %s", cudaGetErrorString(e)); return -6; }
        }
        e = cudaMemsetAsync(tc.counter, 0, sizeof(int), (cudaStream_t)stream);
        if (e != cudaSuccess) { snprintf(err, errlen, "cudaMemsetAsync(tile counter): %s", cudaGetErrorString(e)); return -7; }
        e = cudaFuncSetAttribute(k_cooc_gemm2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES);
        if (e != cudaSuccess) { snprintf(err, errlen, "cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return -7; }
        const int grid = tc.count < n_sm ? tc.count : n_sm;          // persistent: one CTA per SM
        k_cooc_gemm2<<<grid, NUM_THREADS, SMEM_BYTES, (cudaStream_t)stream>>>(tmA, tmB, C, a_rows, b_rows, ldc, k_pad / BK, tc.d, tc.count,
                                                                              symmetric, tc.counter);
    }
    e = cudaGetLastError();
    if (e != cudaSuccess) { snprintf(err, errlen, "k_cooc_gemm launch: %s", cudaGetErrorString(e)); return -7; }
    return 0;
}

extern "C" int fyi_cooc_gemm_launch(const uint8_t* Bt, int n_items, int k_pad, int32_t* C, int ldc, void* stream,
                                    char* err, size_t errlen) {
    return fyi_gemm_u8_nt(Bt, n_items, Bt, n_items, k_pad, C, ldc, 1, stream, err, errlen);
}
