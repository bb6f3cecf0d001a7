// gcx_i8gemm.cuh -- hand-written tcgen05 int8 x int8 -> int32 product for the correlation path.
//
//     T[i][j] = sum_k L[i][k] * R[j][k]            L, R: mb x mb int8, row-major (K contiguous: both operands "K-major")
//
// replaces the cublasGemmEx(CUDA_R_8I) calls of round 1 under the pair sums of lib/corrMatrixCoreSRC.c:8-61 (counts N = I I^T,
// Sx = sum_t 128^t S_t I^T, Sxy = sum 128^(a+b) S_a S_b^T with exact 7-bit slices: every product is exact in int32).
//
// One persistent CTA per SM, 6 warps, three pipelines (the canonical Blackwell GEMM shape):
//   warp 0  TMA producer : cp.async.bulk.tensor.2d loads of a 128 x 128 B tile of L and a 256 x 128 B tile of R per k-block into a
//                          4-stage ring of 128-byte-swizzled shared memory (48 KB per stage), full/empty mbarriers
//   warp 1  MMA issuer   : one elected lane issues tcgen05.mma.cta_group::1.kind::i8, M = 128, N = 256, K = 32 (4 per k-block) with
//                          shared-memory matrix descriptors (SWIZZLE_128B, K-major); accumulators live in TMEM -- 2 x 256 columns of
//                          int32, double buffered so the epilogue of tile t overlaps the products of tile t + 1;
//                          tcgen05.commit frees the smem stage / publishes the accumulator
//   warps 2-5 epilogue   : tcgen05.ld 32x32b.x32 (each warp its 32-lane quarter of TMEM, a thread = one row of the tile) ->
//                          128-bit stores of the int32 row segment
// Descriptor layouts follow cute/arch/mma_sm100_desc.hpp (SmemDescriptor, InstrDescriptor) of the CUTLASS headers in this image.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace gcx {

constexpr int I8_BM = 128, I8_BN = 256, I8_BK = 128;      // tile: rows of L, rows of R, bytes of K per stage
constexpr int I8_STAGES = 4;
constexpr int I8_THREADS = 192;
constexpr int I8_STAGE_BYTES = (I8_BM + I8_BN) * I8_BK;   // 48 KB
constexpr int I8_SMEM = I8_STAGES * I8_STAGE_BYTES + 1024;
constexpr int i8_smem_bytes(int bn, int stages) { return stages * (I8_BM + bn) * I8_BK + 1024; }

__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* tm, uint32_t bar, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst), "l"(tm),
                 "r"(bar), "r"(c0), "r"(c1)
                 : "memory");
}
__device__ __forceinline__ void mbar_init_u32(uint32_t bar, unsigned int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx_u32(uint32_t bar, unsigned int bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive_u32(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait_u32(uint32_t bar, unsigned int parity) {
    unsigned int ok = 0, spins = 0;
    unsigned long long t0 = 0;
    while (!ok) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok)
                     : "r"(bar), "r"(parity)
                     : "memory");
        if (!ok && (++spins & 0xfffu) == 0) {          // a broken pipeline must fail, not hang the device: trap after 4 s
            unsigned long long now;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
            if (t0 == 0) t0 = now;
            else if (now - t0 > 4000000000ull) __trap();
        }
    }
}
// shared-memory matrix descriptor, K-major operand in the 128-byte swizzle: rows of 128 B, 8-row groups 1024 B apart
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr) {
    return (uint64_t)((smem_addr & 0x3ffffu) >> 4)          // start address, 16-byte units
           | ((uint64_t)1 << 16)                            // leading byte offset (unused for swizzled K-major): 1
           | ((uint64_t)(1024 >> 4) << 32)                  // stride byte offset: 8 rows x 128 B
           | ((uint64_t)1 << 46)                            // descriptor version (sm_100)
           | ((uint64_t)2 << 61);                           // layout type SWIZZLE_128B
}
// instruction descriptor: D int32, A / B signed int8, both K-major, M x N
__host__ __device__ constexpr uint32_t umma_idesc_i8(int m, int n) {
    return (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}
__device__ __forceinline__ void umma_i8(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    const uint32_t z = 0;
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %6, %7, %8}, p;\n\t}" ::"r"(tmem_d),
        "l"(da), "l"(db), "r"(idesc), "r"(accumulate), "r"(z), "r"(z), "r"(z), "r"(z)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

// Tile order: the persistent CTAs take consecutive tile numbers, so the tiles in flight at any moment are ~SM-count consecutive
// numbers.  Numbering them down one column of the output made them share one panel of R but read a different panel of L each:
// every tile streamed its own 6 MB of L from HBM (60 GB per product at m = 24,832: the product was DRAM-bound).  Tiles are
// numbered in super-columns of I8_GW column blocks instead -- ~9 row blocks x 8 column blocks in flight, every operand slice
// fetched from HBM once and re-used 8-9 times out of L2.
constexpr int I8_GW = 8;
__device__ __forceinline__ void i8_tile_coords(long long t, int MT, int NT, int& mblk, int& nblk) {
    const long long per_group = (long long)I8_GW * MT;
    const int g = (int)(t / per_group);
    const long long r = t - (long long)g * per_group;
    const int gw = min(I8_GW, NT - g * I8_GW);
    mblk = (int)(r / gw);
    nblk = g * I8_GW + (int)(r - (long long)mblk * gw);
}

// symmetric != 0: L == R, only tiles that touch the lower triangle (i >= j) are computed (the rest of T keeps its old contents)
template <int I8_BN, int I8_STAGES>
__global__ void __launch_bounds__(I8_THREADS, 1)
k_i8gemm_nt(const __grid_constant__ CUtensorMap tmL, const __grid_constant__ CUtensorMap tmR, int* __restrict__ T, int mb, int symmetric) {
    constexpr int I8_STAGE_BYTES = (I8_BM + I8_BN) * I8_BK;
    extern __shared__ unsigned char i8_smem_raw[];
    __shared__ __align__(8) uint64_t bars[2 * I8_STAGES + 4];
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t smem0 = ((uint32_t)__cvta_generic_to_shared(i8_smem_raw) + 1023u) & ~1023u;      // the swizzle needs 1024-byte alignment
    const uint32_t bar0 = (uint32_t)__cvta_generic_to_shared(bars);
    auto full = [&](int s) { return bar0 + 8u * s; };
    auto empty = [&](int s) { return bar0 + 8u * (I8_STAGES + s); };
    auto tfull = [&](int a) { return bar0 + 8u * (2 * I8_STAGES + a); };
    auto tempty = [&](int a) { return bar0 + 8u * (2 * I8_STAGES + 2 + a); };
    const int MT = mb / I8_BM, NT = mb / I8_BN, KB = mb / I8_BK;
    const long long ntiles = (long long)MT * NT;

    if (tid == 0) {
        for (int s = 0; s < I8_STAGES; ++s) { mbar_init_u32(full(s), 1); mbar_init_u32(empty(s), 1); }
        for (int a = 0; a < 2; ++a) { mbar_init_u32(tfull(a), 1); mbar_init_u32(tempty(a), 4); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        const uint32_t dst = (uint32_t)__cvta_generic_to_shared(&tmem_base_s);
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(dst) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = tmem_base_s;

    // symmetric product (L == R): only tiles that touch the LOWER triangle (i >= j) are computed -- the triangle the consumers of
    // the correlation path read (k_corr_finish, k_corr_faccum_sxy); the rest of T keeps its old contents
    auto wanted = [&](long long t, int& mblk, int& nblk) {
        i8_tile_coords(t, MT, NT, mblk, nblk);
        return !symmetric || mblk * I8_BM + (I8_BM - 1) >= nblk * I8_BN;
    };

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            int stage = 0;
            unsigned int phase = 0;
            for (long long t = blockIdx.x; t < ntiles; t += gridDim.x) {
                int mblk, nblk;
                if (!wanted(t, mblk, nblk)) continue;
                for (int kb = 0; kb < KB; ++kb) {
                    mbar_wait_u32(empty(stage), phase ^ 1u);
                    const uint32_t sa = smem0 + (uint32_t)stage * I8_STAGE_BYTES, sb = sa + I8_BM * I8_BK;
                    mbar_expect_tx_u32(full(stage), I8_STAGE_BYTES);
                    tma_load_2d(sa, &tmL, full(stage), kb * I8_BK, mblk * I8_BM);
                    tma_load_2d(sb, &tmR, full(stage), kb * I8_BK, nblk * I8_BN);
                    if (++stage == I8_STAGES) { stage = 0; phase ^= 1u; }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (lane == 0) {
            const uint32_t idesc = umma_idesc_i8(I8_BM, I8_BN);
            int stage = 0;
            unsigned int phase = 0;
            int it = 0;
            for (long long t = blockIdx.x; t < ntiles; t += gridDim.x) {
                int mblk, nblk;
                if (!wanted(t, mblk, nblk)) continue;
                const int as = it & 1;
                const unsigned int aphase = (unsigned int)(it >> 1) & 1u;
                mbar_wait_u32(tempty(as), aphase ^ 1u);                    // the epilogue has drained this accumulator
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t d = tmem_base + (uint32_t)as * I8_BN;
                for (int kb = 0; kb < KB; ++kb) {
                    mbar_wait_u32(full(stage), phase);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint32_t sa = smem0 + (uint32_t)stage * I8_STAGE_BYTES, sb = sa + I8_BM * I8_BK;
#pragma unroll
                    for (int k = 0; k < I8_BK / 32; ++k)
                        umma_i8(d, umma_desc_sw128(sa + 32u * k), umma_desc_sw128(sb + 32u * k), idesc, (kb | k) ? 1u : 0u);
                    umma_commit(empty(stage));                             // frees the stage once these products have read it
                    if (++stage == I8_STAGES) { stage = 0; phase ^= 1u; }
                }
                umma_commit(tfull(as));                                    // accumulator complete
                ++it;
            }
        }
    } else {
        // ===================== epilogue: TMEM -> registers -> global =====================
        const int q = warp & 3;                                             // this warp's 32-lane quarter of TMEM
        int it = 0;
        for (long long t = blockIdx.x; t < ntiles; t += gridDim.x) {
            int mblk, nblk;
            if (!wanted(t, mblk, nblk)) continue;
            const int as = it & 1;
            const unsigned int aphase = (unsigned int)(it >> 1) & 1u;
            mbar_wait_u32(tfull(as), aphase);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const int row = mblk * I8_BM + q * 32 + lane;
            int4* dst = reinterpret_cast<int4*>(T + (long long)row * mb + (long long)nblk * I8_BN);
#pragma unroll 1
            for (int c = 0; c < I8_BN / 32; ++c) {
                uint32_t r[32];
                const uint32_t taddr = tmem_base + (uint32_t)as * I8_BN + (uint32_t)c * 32u + ((uint32_t)(q * 32) << 16);
                asm volatile(
                    "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,"
                    "%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                    : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
                      "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
                      "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
                      "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                    : "r"(taddr)
                    : "memory");
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                for (int v = 0; v < 8; ++v)
                    dst[c * 8 + v] = make_int4((int)r[4 * v], (int)r[4 * v + 1], (int)r[4 * v + 2], (int)r[4 * v + 3]);
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive_u32(tempty(as));
            ++it;
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 2) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
}

// ---------------------------------------------------------------------------------------------------
// CTA-pair variant (cta_group::2): two CTAs of a 2-cluster compute one 256 x 256 tile.  Each CTA stages HALF of both operands
// (its 128 rows of L, its 128 rows of R: 32 KB per k-block instead of 48 KB -> 6 stages, and a third less L2 -> SM traffic per
// multiply-add); the leader's single thread issues tcgen05.mma.cta_group::2 with M = 256, reading both CTAs' shared memory; rows
// 0-127 of the accumulator live in the leader's TMEM, rows 128-255 in the peer's.  Barriers: the TMA loads of both CTAs complete
// on the LEADER's full barrier (peer bit of the mbarrier address cleared, as in cute::SM100_TMA_2SM_LOAD); tcgen05.commit
// multicasts the "stage free" / "accumulator ready" arrivals to both CTAs; the peer's epilogue warps arrive remotely (mapa) on
// the leader's "accumulator drained" barrier.
// ---------------------------------------------------------------------------------------------------
// What the epilogue does with the int32 tile.  mode 0: store it (T, pitch mb).  mode 1: fused accumulation of the correlation path --
// dst[i][j] (= | +=) ldexp((double) t_ij, rhi[i] + (col_exp ? rhi[j] : 0) - shift) for i, j < m (dst: float64 m x m; a symmetric product
// only touches j <= i) -- the arithmetic of k_corr_accum32 / k_corr_faccum / k_corr_faccum_sxy(sym = 0) without the int32 round trip.
struct I8Epilogue {
    int mode;
    int m;
    double* dst;
    const int* rhi;       // per-row exponents (nullptr: all zero)
    int shift, first, col_exp;
};

constexpr int I8P_STAGES = 6;
constexpr int I8P_STAGE_BYTES = 2 * 128 * I8_BK;            // 32 KB per CTA per k-block
constexpr int I8P_SMEM = I8P_STAGES * I8P_STAGE_BYTES + 1024;

__device__ __forceinline__ void tma_load_2d_pair(uint32_t dst, const CUtensorMap* tm, uint32_t leader_bar, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
                 "l"(tm), "r"(leader_bar), "r"(c0), "r"(c1)
                 : "memory");
}
__device__ __forceinline__ void umma_i8_pair(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    const uint32_t z = 0;
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::i8 [%0], %1, %2, %3, {%5, %6, %7, %8, %9, %10, %11, %12}, p;\n\t}" ::"r"(
            tmem_d),
        "l"(da), "l"(db), "r"(idesc), "r"(accumulate), "r"(z), "r"(z), "r"(z), "r"(z), "r"(z), "r"(z), "r"(z), "r"(z)
        : "memory");
}
__device__ __forceinline__ void umma_commit_pair(uint32_t bar) {          // arrives on `bar` in BOTH CTAs of the pair
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar), "h"((uint16_t)3)
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive_on_cta(uint32_t bar, uint32_t cta_rank) {
    asm volatile("{\n\t.reg .b32 ra;\n\tmapa.shared::cluster.u32 ra, %0, %1;\n\tmbarrier.arrive.release.cluster.shared::cluster.b64 _, [ra];\n\t}" ::"r"(bar),
                 "r"(cta_rank)
                 : "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(I8_THREADS, 1)
k_i8gemm_nt_pair(const __grid_constant__ CUtensorMap tmL, const __grid_constant__ CUtensorMap tmR, int* __restrict__ T, int mb, int symmetric,
                 I8Epilogue E) {
    extern __shared__ unsigned char i8_smem_raw[];
    __shared__ __align__(8) uint64_t bars[2 * I8P_STAGES + 4];
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    uint32_t rank;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
    const bool leader = rank == 0;
    const uint32_t smem0 = ((uint32_t)__cvta_generic_to_shared(i8_smem_raw) + 1023u) & ~1023u;
    const uint32_t bar0 = (uint32_t)__cvta_generic_to_shared(bars);
    auto full = [&](int s) { return bar0 + 8u * s; };
    auto empty = [&](int s) { return bar0 + 8u * (I8P_STAGES + s); };
    auto tfull = [&](int a) { return bar0 + 8u * (2 * I8P_STAGES + a); };
    auto tempty = [&](int a) { return bar0 + 8u * (2 * I8P_STAGES + 2 + a); };
    const int MT = mb / 256, KB = mb / I8_BK;
    const long long ntiles = (long long)MT * MT;
    const int pair = blockIdx.x >> 1, npairs = gridDim.x >> 1;

    if (tid == 0) {
        for (int s = 0; s < I8P_STAGES; ++s) { mbar_init_u32(full(s), 1); mbar_init_u32(empty(s), 1); }
        for (int a = 0; a < 2; ++a) { mbar_init_u32(tfull(a), 1); mbar_init_u32(tempty(a), 8); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        const uint32_t dst = (uint32_t)__cvta_generic_to_shared(&tmem_base_s);
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(dst) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    cluster_sync_all();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = tmem_base_s;

    auto wanted = [&](long long t, int& mblk, int& nblk) {
        i8_tile_coords(t, MT, MT, mblk, nblk);
        return !symmetric || mblk >= nblk;          // 256 x 256 tiles: the ones that touch i >= j
    };

    if (warp == 0) {
        // ===================== TMA producer (both CTAs: each its half of L's and of R's rows) =====================
        if (lane == 0) {
            int stage = 0;
            unsigned int phase = 0;
            for (long long t = pair; t < ntiles; t += npairs) {
                int mblk, nblk;
                if (!wanted(t, mblk, nblk)) continue;
                for (int kb = 0; kb < KB; ++kb) {
                    mbar_wait_u32(empty(stage), phase ^ 1u);
                    const uint32_t sa = smem0 + (uint32_t)stage * I8P_STAGE_BYTES, sb = sa + 128 * I8_BK;
                    const uint32_t lbar = full(stage) & 0xFEFFFFFFu;                 // the leader's barrier collects both CTAs' bytes
                    if (leader) mbar_expect_tx_u32(full(stage), 2 * I8P_STAGE_BYTES);
                    tma_load_2d_pair(sa, &tmL, lbar, kb * I8_BK, mblk * 256 + (int)rank * 128);
                    tma_load_2d_pair(sb, &tmR, lbar, kb * I8_BK, nblk * 256 + (int)rank * 128);
                    if (++stage == I8P_STAGES) { stage = 0; phase ^= 1u; }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (leader CTA only) =====================
        if (leader && lane == 0) {
            const uint32_t idesc = umma_idesc_i8(256, 256);
            int stage = 0;
            unsigned int phase = 0;
            int it = 0;
            for (long long t = pair; t < ntiles; t += npairs) {
                int mblk, nblk;
                if (!wanted(t, mblk, nblk)) continue;
                const int as = it & 1;
                const unsigned int aphase = (unsigned int)(it >> 1) & 1u;
                mbar_wait_u32(tempty(as), aphase ^ 1u);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t d = tmem_base + (uint32_t)as * 256u;
                for (int kb = 0; kb < KB; ++kb) {
                    mbar_wait_u32(full(stage), phase);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint32_t sa = smem0 + (uint32_t)stage * I8P_STAGE_BYTES, sb = sa + 128 * I8_BK;
#pragma unroll
                    for (int k = 0; k < I8_BK / 32; ++k)
                        umma_i8_pair(d, umma_desc_sw128(sa + 32u * k), umma_desc_sw128(sb + 32u * k), idesc, (kb | k) ? 1u : 0u);
                    umma_commit_pair(empty(stage));
                    if (++stage == I8P_STAGES) { stage = 0; phase ^= 1u; }
                }
                umma_commit_pair(tfull(as));
                ++it;
            }
        }
    } else {
        // ===================== epilogue (both CTAs: each its 128 rows of the tile) =====================
        const int q = warp & 3;
        int it = 0;
        for (long long t = pair; t < ntiles; t += npairs) {
            int mblk, nblk;
            if (!wanted(t, mblk, nblk)) continue;
            const int as = it & 1;
            const unsigned int aphase = (unsigned int)(it >> 1) & 1u;
            mbar_wait_u32(tfull(as), aphase);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const int row = mblk * 256 + (int)rank * 128 + q * 32 + lane;
            int4* dst = reinterpret_cast<int4*>(T + (long long)row * mb + (long long)nblk * 256);
#pragma unroll 1
            for (int c = 0; c < 256 / 32; ++c) {
                uint32_t r[32];
                const uint32_t taddr = tmem_base + (uint32_t)as * 256u + (uint32_t)c * 32u + ((uint32_t)(q * 32) << 16);
                asm volatile(
                    "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,"
                    "%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                    : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
                      "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
                      "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
                      "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                    : "r"(taddr)
                    : "memory");
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                if (E.mode == 0) {
#pragma unroll
                    for (int v = 0; v < 8; ++v)
                        dst[c * 8 + v] = make_int4((int)r[4 * v], (int)r[4 * v + 1], (int)r[4 * v + 2], (int)r[4 * v + 3]);
                } else if (row < E.m) {
                    const int j0 = nblk * 256 + c * 32;
                    const int ei = (E.rhi != nullptr ? E.rhi[row] : 0) - E.shift;
                    double* d = E.dst + (long long)row * E.m + j0;
                    const int jend = min(32, (symmetric ? min(E.m, row + 1) : E.m) - j0);
#pragma unroll
                    for (int v = 0; v < 32; ++v) {
                        if (v < jend) {
                            const int e = ei + (E.col_exp ? E.rhi[j0 + v] : 0);
                            const double val = ldexp((double)(int)r[v], e);
                            d[v] = E.first ? val : d[v] + val;
                        }
                    }
                }
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive_on_cta(tempty(as), 0u);          // the leader's MMA issuer counts 8 warps (4 per CTA)
            ++it;
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    cluster_sync_all();
    if (warp == 2) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
}

}  // namespace gcx
