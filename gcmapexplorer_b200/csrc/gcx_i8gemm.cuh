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
    unsigned int ok = 0;
    while (!ok) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok)
                     : "r"(bar), "r"(parity)
                     : "memory");
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

// symmetric != 0: L == R, only tiles that touch the upper triangle (i <= j) are computed (the rest of T keeps its old contents)
__global__ void __launch_bounds__(I8_THREADS, 1)
k_i8gemm_nt(const __grid_constant__ CUtensorMap tmL, const __grid_constant__ CUtensorMap tmR, int* __restrict__ T, int mb, int symmetric) {
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

    // a tile is skipped when it lies entirely below the diagonal of a symmetric product
    auto wanted = [&](long long t, int& mblk, int& nblk) {
        mblk = (int)(t % MT);
        nblk = (int)(t / MT);
        return !symmetric || mblk * I8_BM <= nblk * I8_BN + (I8_BN - 1);
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

}  // namespace gcx
