// gcx_kernels.cuh -- hand-written sm_100a kernels for the contact-map normalization hot path.
//
// Data layout in HBM: the rank's row block of the dense float32 n x n map, row-major, row pitch `ld`
// floats (ld % 32 == 0, so every row starts on a 128-byte line and 128-bit loads are always aligned;
// pad columns are zero).  All O(n) vectors are float64, length >= ld, zero padded.
//
// Every matrix pass uses the same decomposition: the block is cut into bands of R rows and chunks of
// MV_THREADS float4 columns; the flattened (band, chunk) unit list is split evenly over a persistent
// grid of (SM count x occupancy) CTAs, so no CTA has more than one unit more than another.  A thread
// keeps the column-vector entries for its float4 in registers and reuses them across the R rows of the
// band, which cuts L2 traffic for the vector to 1/R of the matrix traffic.
#pragma once
#include <cooperative_groups.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace gcx {

constexpr int MV_THREADS = 256;   // threads per CTA for matrix passes
constexpr int EP_THREADS = 1024;  // threads per CTA of the vector-step kernels
constexpr int EP_CL = 8;          // CTAs per thread-block cluster running a vector step (8 SMs, DSMEM reductions)

// ---------------------------------------------------------------------------------------------------
// small device helpers
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ float4 ld_stream_f4(const float4* p) {
    // streaming 128-bit load: read-only path, do not allocate in L1 (the map is read once per pass)
    float4 r;
    asm("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
        : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
        : "l"(p));
    return r;
}

__device__ __forceinline__ void st_stream_f4(float4* p, const float4& v) {
    asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
                 : "memory");
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_min(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmin(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ long long warp_sum_ll(long long v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Deterministic block reductions (fixed tree).  `sm` holds >= 32 doubles.  All threads get the result.
enum { RED_SUM = 0, RED_MIN = 1, RED_MAX = 2 };
template <int OP>
__device__ __forceinline__ double block_reduce(double v, double* sm) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    v = OP == RED_SUM ? warp_sum(v) : (OP == RED_MIN ? warp_min(v) : warp_max(v));
    if (lane == 0) sm[warp] = v;
    __syncthreads();
    if (warp == 0) {
        const double ident = OP == RED_SUM ? 0.0 : (OP == RED_MIN ? __longlong_as_double(0x7ff0000000000000LL)
                                                                   : __longlong_as_double(0xfff0000000000000LL));
        double x = lane < nw ? sm[lane] : ident;
        x = OP == RED_SUM ? warp_sum(x) : (OP == RED_MIN ? warp_min(x) : warp_max(x));
        if (lane == 0) sm[0] = x;
    }
    __syncthreads();
    const double r = sm[0];
    __syncthreads();
    return r;
}

// ---------------------------------------------------------------------------------------------------
// Team: one thread-block cluster (EP_CL CTAs x EP_THREADS threads) that runs a vector step.  Element i is
// always handled by the same thread (i = gtid + k*nthreads), so successive elementwise phases need no
// synchronisation; only reductions do.  A reduction is: fixed-tree block reduce -> one slot in this CTA's
// shared memory -> cluster barrier -> warp 0 reads the EP_CL slots through distributed shared memory and
// folds them in rank order.  Every CTA folds the same values in the same order, so all CTAs (and all ranks
// running the same step on the same gathered vector) hold bitwise identical scalars.
// ---------------------------------------------------------------------------------------------------
struct Team {
    cooperative_groups::cluster_group cl;
    double* slots;   // shared, [2][4]: double-buffered so one cluster barrier per reduction is enough
    double* sm;      // shared, [32]: block-reduce scratch
    int par;
    unsigned int nb, gtid, nthreads;
    __device__ Team(double* slots_, double* sm_)
        : cl(cooperative_groups::this_cluster()), slots(slots_), sm(sm_), par(0) {
        nb = cl.num_blocks();
        gtid = cl.block_rank() * blockDim.x + threadIdx.x;
        nthreads = nb * blockDim.x;
    }
    template <int OP, int K>
    __device__ __forceinline__ void reduce(double (&v)[K]) {
        static_assert(K <= 4, "at most 4 values per reduction");
#pragma unroll
        for (int k = 0; k < K; ++k) v[k] = block_reduce<OP>(v[k], sm);
        if (threadIdx.x == 0) {
#pragma unroll
            for (int k = 0; k < K; ++k) slots[par * 4 + k] = v[k];
        }
        cl.sync();
        if (threadIdx.x < 32) {
            const unsigned int lane = threadIdx.x;
#pragma unroll
            for (int k = 0; k < K; ++k) {
                double x = 0.0;
                if (lane < nb) x = *cl.map_shared_rank(&slots[par * 4 + k], lane);
                double acc = __shfl_sync(0xffffffffu, x, 0);
                for (unsigned int r = 1; r < nb; ++r) {
                    const double y = __shfl_sync(0xffffffffu, x, r);
                    acc = OP == RED_SUM ? acc + y : (OP == RED_MIN ? fmin(acc, y) : fmax(acc, y));
                }
                if (lane == 0) sm[k] = acc;
            }
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < K; ++k) v[k] = sm[k];
        __syncthreads();
        par ^= 1;
    }
    template <int OP>
    __device__ __forceinline__ double reduce1(double v) {
        double a[1] = {v};
        reduce<OP, 1>(a);
        return a[0];
    }
    // no CTA may exit while a peer can still read its shared memory
    __device__ __forceinline__ void finish() { cl.sync(); }
};

// order-preserving float <-> uint32 encoding for atomicMin/atomicMax on floats of either sign
__device__ __forceinline__ unsigned int f2ord(float f) {
    unsigned int u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__host__ __device__ __forceinline__ float ord2f(unsigned int u) {
    unsigned int b = (u & 0x80000000u) ? (u & 0x7fffffffu) : ~u;
#ifdef __CUDA_ARCH__
    return __uint_as_float(b);
#else
    float f;
    memcpy(&f, &b, 4);
    return f;
#endif
}

__host__ __device__ __forceinline__ uint64_t mix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}

// Even split of U work units over G CTAs: CTA c owns [c*U/G, (c+1)*U/G).
__device__ __forceinline__ long long unit_begin(long long c, long long U, long long G) { return c * U / G; }
// ---------------------------------------------------------------------------------------------------
// K2/K3: q[row0 + r] = sum_j A[r][j] * z[j]   (float32 map, float64 vector, float64 accumulation)
// Replaces np.sum(matrix, axis=0) of lib/normalizeCore.pyx:42 (symmetric map: column sums = row sums
// of the bias-scaled map) and np.dot(A, x) of lib/normalizeKnightRuiz.pyx:120, 139, 169.
// Algorithmic HBM bytes per launch: 4 * nloc * n (one read of the block).
// Bands that straddle two CTAs' unit ranges leave partial sums in `ws`; the last CTA to finish folds them
// in CTA order, so the result is bitwise reproducible for a given grid size.
// ---------------------------------------------------------------------------------------------------
template <int R, int OCC>
__global__ void __launch_bounds__(MV_THREADS, OCC)
k_matvec(const float* __restrict__ A, long long ld, int nloc, int n4, const double* __restrict__ z,
         double* __restrict__ q, double* ws, unsigned int* ticket, const int* __restrict__ done) {
    if (done != nullptr && *done) return;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nch = (n4 + MV_THREADS - 1) / MV_THREADS;
    const int nb = (nloc + R - 1) / R;
    const long long U = (long long)nb * nch, G = gridDim.x, c = blockIdx.x;
    long long u = unit_begin(c, U, G);
    const long long u1 = unit_begin(c + 1, U, G);
    const int bfirst = (int)(u / nch);
    __shared__ double red[MV_THREADS / 32][R];
    const double2* __restrict__ z2 = reinterpret_cast<const double2*>(z);

    while (u < u1) {
        const int b = (int)(u / nch);
        const int k0 = (int)(u - (long long)b * nch);
        const long long rem = u1 - u;
        const int k1 = (int)(((long long)(nch - k0) < rem) ? nch : (k0 + rem));
        const float4* rowp[R];
#pragma unroll
        for (int rr = 0; rr < R; ++rr) {
            int r = b * R + rr;
            r = r < nloc ? r : nloc - 1;
            rowp[rr] = reinterpret_cast<const float4*>(A + (long long)r * ld);
        }
        double acc[R];
#pragma unroll
        for (int rr = 0; rr < R; ++rr) acc[rr] = 0.0;

        for (int k = k0; k < k1; ++k) {
            const int c4 = k * MV_THREADS + tid;
            if (c4 < n4) {
                float4 a[R];
#pragma unroll
                for (int rr = 0; rr < R; ++rr) a[rr] = ld_stream_f4(rowp[rr] + c4);
                const double2 za = __ldg(z2 + 2 * c4), zb = __ldg(z2 + 2 * c4 + 1);
#pragma unroll
                for (int rr = 0; rr < R; ++rr) {
                    acc[rr] = fma((double)a[rr].x, za.x, acc[rr]);
                    acc[rr] = fma((double)a[rr].y, za.y, acc[rr]);
                    acc[rr] = fma((double)a[rr].z, zb.x, acc[rr]);
                    acc[rr] = fma((double)a[rr].w, zb.y, acc[rr]);
                }
            }
        }
#pragma unroll
        for (int rr = 0; rr < R; ++rr) acc[rr] = warp_sum(acc[rr]);
        if (lane == 0) {
#pragma unroll
            for (int rr = 0; rr < R; ++rr) red[warp][rr] = acc[rr];
        }
        __syncthreads();
        if (tid < R) {
            double s = 0.0;
#pragma unroll
            for (int w = 0; w < MV_THREADS / 32; ++w) s += red[w][tid];
            const int r = b * R + tid;
            if (k0 == 0 && k1 == nch) {
                if (r < nloc) q[r] = s;
            } else {
                ws[((long long)c * 2 + (b == bfirst ? 0 : 1)) * R + tid] = s;
            }
        }
        __syncthreads();
        u += (k1 - k0);
    }

    // ---- last-arriving CTA folds the split bands ----
    __shared__ int s_last;
    __threadfence();
    __syncthreads();
    if (tid == 0) s_last = (atomicAdd(ticket, 1u) == (unsigned int)(G - 1));
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    // One thread per CTA boundary (G-1 of them) instead of one per band: a band is split only where a
    // boundary falls strictly inside it.  The thread of the first such boundary folds the partials of
    // every CTA that touched the band, in CTA order.
    for (long long cb = 1 + tid; cb < G; cb += MV_THREADS) {
        const long long ub = unit_begin(cb, U, G);
        const int b = (int)(ub / nch);
        const long long band0 = (long long)b * nch, band1 = band0 + nch;
        if (ub == band0 || ub >= U) continue;                       // boundary on a band edge
        const long long ubp = unit_begin(cb - 1, U, G);
        if (ubp > band0) continue;                                   // an earlier boundary owns this band
        double s[R];
        {
            const int slot = (b == (int)(ubp / nch)) ? 0 : 1;        // CTA cb-1: band b is its last band
#pragma unroll
            for (int rr = 0; rr < R; ++rr) s[rr] = __ldcg(ws + ((cb - 1) * 2 + slot) * R + rr);
        }
        long long cc = cb, ucc = ub;
        while (cc < G && ucc < band1) {
            const long long unext = unit_begin(cc + 1, U, G);
            if (unext > ucc) {                                       // non-empty CTA: band b is its first band
#pragma unroll
                for (int rr = 0; rr < R; ++rr) s[rr] += __ldcg(ws + (cc * 2) * R + rr);
            }
            ++cc;
            ucc = unext;
        }
#pragma unroll
        for (int rr = 0; rr < R; ++rr) {
            const int r = b * R + rr;
            if (r < nloc) q[r] = s[rr];
        }
    }
    if (tid == 0) *ticket = 0u;
}

// ---------------------------------------------------------------------------------------------------
// K2/K3, TMA-bulk variant: the same (band, chunk) decomposition and arithmetic as k_matvec, but the map is
// staged through shared memory by the bulk async-copy engine (cp.async.bulk = SASS UBLKCP) instead of
// per-thread LDG.128: one producer lane keeps STAGES x R row segments (R x 4 KB per stage) in flight behind
// mbarriers, 8 consumer warps read them back with conflict-free LDS.128.  Bytes in flight no longer depend on
// registers or occupancy.  Selected with GCX_MATVEC_VARIANT >= 7; see profiles/ for the measured comparison.
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned int smem_u32(const void* p) { return (unsigned int)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, unsigned int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, unsigned int bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, unsigned int parity) {
    unsigned int ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned int parity) {
    while (!mbar_try_wait(bar, parity)) {
    }
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned int bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

__device__ __forceinline__ unsigned long long globaltimer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}

// The map is read once per pass and never reused; an evict-first L2 policy on its bulk copies keeps the few MB of row / column
// partials, vectors and tables that the NEXT launch's vector phase reads resident in L2 instead of being flushed to HBM by the
// 1.3 GB stream in between.
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ void bulk_g2s_hint(void* dst, const void* src, unsigned int bytes, uint64_t* bar, uint64_t policy) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
                 : "memory");
}

constexpr int TMA_CONSUMERS = MV_THREADS;          // 8 consumer warps
constexpr int TMA_THREADS = MV_THREADS + 32;       // + 1 producer warp

// Split bands: one thread per CTA boundary folds the partial sums of every CTA that touched the band, in CTA
// order (bitwise reproducible for a given grid).  Called by the last CTA to finish.
template <int R>
__device__ __forceinline__ void fold_split_bands(double* q, const double* ws, long long U, long long G, int nch, int nloc, int nthreads) {
    for (long long cb = 1 + threadIdx.x; cb < G; cb += nthreads) {
        const long long ub = unit_begin(cb, U, G);
        const int b = (int)(ub / nch);
        const long long band0 = (long long)b * nch, band1 = band0 + nch;
        if (ub == band0 || ub >= U) continue;                        // boundary on a band edge
        const long long ubp = unit_begin(cb - 1, U, G);
        if (ubp > band0) continue;                                   // an earlier boundary owns this band
        double s[R];
        {
            const int slot = (b == (int)(ubp / nch)) ? 0 : 1;        // CTA cb-1: band b is its last band
#pragma unroll
            for (int rr = 0; rr < R; ++rr) s[rr] = __ldcg(ws + ((cb - 1) * 2 + slot) * R + rr);
        }
        long long cc = cb, ucc = ub;
        while (cc < G && ucc < band1) {
            const long long unext = unit_begin(cc + 1, U, G);
            if (unext > ucc) {                                       // non-empty CTA: band b is its first band
#pragma unroll
                for (int rr = 0; rr < R; ++rr) s[rr] += __ldcg(ws + (cc * 2) * R + rr);
            }
            ++cc;
            ucc = unext;
        }
#pragma unroll
        for (int rr = 0; rr < R; ++rr) {
            const int r = b * R + rr;
            if (r < nloc) q[r] = s[rr];
        }
    }
}

// The producer/consumer main loop.  ZNC: read z through the non-coherent path (allowed only when this kernel
// never writes z).  Ends with every thread past a CTA-wide barrier.
template <int R, int STAGES, bool ZNC>
__device__ __forceinline__ void matvec_tma_core(const float* __restrict__ A, long long ld, int nloc, int n4, const double* z, double* q,
                                                double* ws, float4* tiles, uint64_t* full_bar, uint64_t* empty_bar,
                                                double (*red)[R]) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nch = (n4 + MV_THREADS - 1) / MV_THREADS;
    const int nb = (nloc + R - 1) / R;
    const long long U = (long long)nb * nch, G = gridDim.x, c = blockIdx.x;
    const long long u0 = unit_begin(c, U, G), u1 = unit_begin(c + 1, U, G);
    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], MV_THREADS / 32);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    if (warp == MV_THREADS / 32) {
        // ---------------- producer: one lane drives the bulk-copy engine ----------------
        if (lane == 0) {
            int stage = 0;
            unsigned int phase = 0;
            for (long long u = u0; u < u1; ++u) {
                mbar_wait(&empty_bar[stage], phase ^ 1u);
                const int b = (int)(u / nch), k = (int)(u - (long long)b * nch);
                const int c0 = k * MV_THREADS;
                const int nf4 = min(MV_THREADS, n4 - c0);
                const unsigned int row_bytes = (unsigned int)nf4 * 16u;
                mbar_expect_tx(&full_bar[stage], row_bytes * R);
#pragma unroll
                for (int rr = 0; rr < R; ++rr) {
                    int r = b * R + rr;
                    r = r < nloc ? r : nloc - 1;
                    bulk_g2s(tiles + ((size_t)stage * R + rr) * MV_THREADS, A + (long long)r * ld + 4LL * c0, row_bytes, &full_bar[stage]);
                }
                if (++stage == STAGES) {
                    stage = 0;
                    phase ^= 1u;
                }
            }
        }
    } else {
        // ---------------- consumers ----------------
        const double2* z2 = reinterpret_cast<const double2*>(z);
        const int bfirst = (int)(u0 / nch);
        int stage = 0;
        unsigned int phase = 0;
        long long u = u0;
        while (u < u1) {
            const int b = (int)(u / nch);
            const int k0 = (int)(u - (long long)b * nch);
            const long long rem = u1 - u;
            const int k1 = (int)(((long long)(nch - k0) < rem) ? nch : (k0 + rem));
            double acc[R];
#pragma unroll
            for (int rr = 0; rr < R; ++rr) acc[rr] = 0.0;
            for (int k = k0; k < k1; ++k) {
                const int c4 = k * MV_THREADS + tid;
                double2 za = make_double2(0.0, 0.0), zb = za;
                if (c4 < n4) {
                    if (ZNC) {
                        za = __ldg(z2 + 2 * c4);
                        zb = __ldg(z2 + 2 * c4 + 1);
                    } else {
                        za = z2[2 * c4];
                        zb = z2[2 * c4 + 1];
                    }
                }
                mbar_wait(&full_bar[stage], phase);
                if (c4 < n4) {
                    float4 a[R];
#pragma unroll
                    for (int rr = 0; rr < R; ++rr) a[rr] = tiles[((size_t)stage * R + rr) * MV_THREADS + tid];
#pragma unroll
                    for (int rr = 0; rr < R; ++rr) {
                        acc[rr] = fma((double)a[rr].x, za.x, acc[rr]);
                        acc[rr] = fma((double)a[rr].y, za.y, acc[rr]);
                        acc[rr] = fma((double)a[rr].z, zb.x, acc[rr]);
                        acc[rr] = fma((double)a[rr].w, zb.y, acc[rr]);
                    }
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&empty_bar[stage]);
                if (++stage == STAGES) {
                    stage = 0;
                    phase ^= 1u;
                }
            }
#pragma unroll
            for (int rr = 0; rr < R; ++rr) acc[rr] = warp_sum(acc[rr]);
            if (lane == 0) {
#pragma unroll
                for (int rr = 0; rr < R; ++rr) red[warp][rr] = acc[rr];
            }
            asm volatile("bar.sync 1, %0;" ::"n"(TMA_CONSUMERS) : "memory");
            if (tid < R) {
                double s = 0.0;
#pragma unroll
                for (int w = 0; w < MV_THREADS / 32; ++w) s += red[w][tid];
                const int r = b * R + tid;
                if (k0 == 0 && k1 == nch) {
                    if (r < nloc) q[r] = s;
                } else {
                    ws[((long long)c * 2 + (b == bfirst ? 0 : 1)) * R + tid] = s;
                }
            }
            asm volatile("bar.sync 1, %0;" ::"n"(TMA_CONSUMERS) : "memory");
            u += (k1 - k0);
        }
    }
    __threadfence();
    __syncthreads();
}

template <int R, int STAGES, int OCC>
__global__ void __launch_bounds__(TMA_THREADS, OCC)
k_matvec_tma(const float* __restrict__ A, long long ld, int nloc, int n4, const double* __restrict__ z, double* __restrict__ q,
             double* ws, unsigned int* ticket, const int* __restrict__ done) {
    if (done != nullptr && *done) return;
    extern __shared__ __align__(128) unsigned char tma_smem[];
    __shared__ __align__(8) uint64_t full_bar[STAGES];
    __shared__ __align__(8) uint64_t empty_bar[STAGES];
    __shared__ double red[MV_THREADS / 32][R];
    __shared__ int s_last;
    matvec_tma_core<R, STAGES, true>(A, ld, nloc, n4, z, q, ws, reinterpret_cast<float4*>(tma_smem), full_bar, empty_bar, red);
    const int nch = (n4 + MV_THREADS - 1) / MV_THREADS;
    const long long U = (long long)((nloc + R - 1) / R) * nch, G = gridDim.x;
    if (threadIdx.x == 0) s_last = (atomicAdd(ticket, 1u) == (unsigned int)(G - 1));
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    fold_split_bands<R>(q, ws, U, G, nch, nloc, TMA_THREADS);
    if (threadIdx.x == 0) *ticket = 0u;
}

// ---------------------------------------------------------------------------------------------------
// K2/K3, symmetric variant: t = A z for a SYMMETRIC map, reading only the upper triangle of the dense
// layout (north_star: "...or as the upper triangle, since the map is symmetric").  HBM bytes per launch:
// ~2 n (n + 1024) instead of 4 n^2.
//
// Units are (panel k of 1024 columns, band b of 4 rows) with rows <= last column of the panel, enumerated
// panel-major so that a CTA's contiguous share stays inside a few panels.  The 8 consumer warps form NG groups
// that take the CTA's units round-robin; a thread of a group owns 4*NG columns of the panel:
//   column part  y_j += A_ij z_i  (j > i)  accumulates in registers across all bands of the panel and is flushed
//                once per (CTA, panel, group) into colpart;
//   row part     x_i += A_ij z_j  (j >= i) is reduced over the warp's columns with a 6-shuffle butterfly and
//                stored as one partial per (panel, warp-in-group, row) in rowpart -- no CTA-wide barrier, no atomics.
// The fold adds, in a fixed order, the row partials of the panels right of i and the column partials of the
// CTAs that walked i's panel.  Everything is deterministic for a given grid.  Tables with every CTA's share
// (SymTab) are computed once per map on the host, so the kernels do no 64-bit divisions.
// ---------------------------------------------------------------------------------------------------
constexpr int SYM_R = 4;
constexpr int SYM_BPP = 1024 / SYM_R;        // bands per 1024 rows
#ifndef GCX_SYM_CFG
#define GCX_SYM_CFG 0
#endif
#if GCX_SYM_CFG == 1                         // one CTA per SM: 16 consumer warps in 4 groups, 12 stages (192 KB)
constexpr int SYM_CW = 16, SYM_NG = 4, SYM_STAGES = 12, SYM_OCC = 1;
#else                                        // two CTAs per SM: 8 consumer warps in 2 groups, 6 stages (96 KB) each
constexpr int SYM_CW = 8, SYM_NG = 2, SYM_STAGES = 6, SYM_OCC = 2;
#endif
constexpr int SYM_THREADS = (SYM_CW + 1) * 32;   // consumer warps + 1 producer warp

// Geometry of one rank's share of the upper triangle, tabulated on the host once per map.  The rank owns the global
// rows [row0, row0 + nloc) (row0 a multiple of 1024 when sharded).  Panel k holds the local bands whose global rows lie
// in [row0, min(row0 + nloc, 1024 (k+1))); panel_u0 is the running count of those bands (units) over the panels.
struct SymTab {
    const long long* panel_u0; // [nch + 1] first unit of every panel (units of panels left of row0's panel: none)
    const long long* cta_u0;   // [G + 1] first unit of every CTA's share
    const int* cta_kf;         // [G] panel of a share's first unit
    const int* cta_kl;         // [G] panel of a share's last unit
    const int* panel_cta;      // [nch] first CTA whose (non-empty) share intersects the panel (G if none)
    const int* panel_cta_hi;   // [nch] last such CTA (-1 if none); every CTA in between has units in the panel
    const int* panel_full;     // [nch] 1: the panel is an off-diagonal block assigned to this rank (every element counts for
                               //      the row AND the column part); 0: the rank's own diagonal region (j >= i / j > i rules)
    int kfold_lo;              // first panel with units on this rank (row-partial fold starts here)
    const int* fold_k;         // sharded: the panels this rank walks (ascending) -- the only row partials that can be non-zero
    int n_fold_k;              // 0: every panel from kfold_lo (or the bin's own panel) on
    long long* dbg_cycles;     // optional [G]: cycles every CTA spent in the main loop (GCX_SYM_DEBUG=1), else nullptr
    long long* dbg_stamps;     // optional [G][16]: globaltimer (ns) at the phases of launch dbg_launch of k_ic_pass_symd
    int dbg_launch;
    int l2_hint;               // 1: evict-first L2 policy on the map's bulk copies (GCX_SYM_L2HINT, default 1)
    // dynamic pool: the units of a few large panels are not part of any CTA's static share; they are cut into chunks of
    // consecutive bands of one panel and handed to whichever CTA finishes its static share first (atomic counter).  A chunk
    // is always walked by one CTA in a fixed order and leaves its own column partial, so results do not depend on who ran it.
    int ndyn;                  // number of chunks
    const int* dyn_k;          // [ndyn] panel of the chunk
    const int* dyn_b0;         // [ndyn] first local band
    const int* dyn_cnt;        // [ndyn] number of bands
    const int* dyn_order;      // [ndyn] hand-out order: ticket -> chunk (large chunks first, the small ones close the pass)
    const int* panel_dyn0;     // [nch + 1] first chunk of every panel
    double* colpart_dyn;       // [ndyn][NG][1024]
    unsigned int* dyn_counter; // zeroed by the host before every launch
    int nslots;                // panels a share may touch (sizes colpart)
    int row0, nloc;            // this rank's global row block
};

// MODE: SYM_ALL = the whole walk; SYM_BEGIN = barrier set-up + the producer's first min(STAGES, u1 - u0) loads only (the tiles of A
// do not depend on the vector, so a fused pass issues them BEFORE its vector phase and grid barrier); SYM_CONT = the rest of a walk
// started with SYM_BEGIN.
enum { SYM_ALL = 0, SYM_BEGIN = 1, SYM_CONT = 2 };
template <int NG, int STAGES, bool ZNC, int MODE = SYM_ALL>
__device__ __forceinline__ void matvec_sym_range(const float* __restrict__ A, long long ld, int n4, const double* z, double* rowpart,
                                                 double* colslots, const SymTab& T, float4* tiles, uint64_t* full_bar, uint64_t* empty_bar,
                                                 double* gred, double* cred, long long u0, long long u1, int kfirst, int klast, bool single_panel) {
    // Walks the units [u0, u1).  Static share: unit numbers of the panel-major list (panel_u0 maps them to (panel, band)),
    // panels kfirst..klast, column partials to colslots[(panel - kfirst)][group].  Dynamic chunk (single_panel): the units are the
    // bands u0..u1-1 of panel kfirst.  Every thread of the CTA must call this; it starts and ends with a CTA-wide barrier.
    constexpr int R = SYM_R;
    constexpr int WPG = SYM_CW / NG;                 // warps per group
    constexpr int TG = WPG * 32;                     // threads per group
    constexpr int NQ = MV_THREADS / TG;              // float4 columns of the 1024-column panel per thread
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nloc = T.nloc, row0 = T.row0;
    const long long NOEND = 0x7fffffffffffffffLL;
    if (MODE != SYM_CONT) {
        if (tid == 0) {
#pragma unroll
            for (int s = 0; s < STAGES; ++s) {
                mbar_init(&full_bar[s], 1);
                mbar_init(&empty_bar[s], WPG);
            }
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncthreads();
    }
    const long long upre = (u1 - u0 < (long long)STAGES) ? u1 : u0 + STAGES;      // end of the part SYM_BEGIN issues
    if (u0 < u1) {
    if (warp == SYM_CW) {
        // ---------------- producer ----------------
        if (lane == 0) {
            const long long ua = MODE == SYM_CONT ? upre : u0, ub = MODE == SYM_BEGIN ? upre : u1;
            int stage = (int)((ua - u0) % STAGES);
            unsigned int phase = (unsigned int)((ua - u0) / STAGES) & 1u;
            int k = kfirst;
            long long pbeg = single_panel ? 0 : T.panel_u0[k];
            long long pend = single_panel ? NOEND : T.panel_u0[k + 1];
            const uint64_t pol = l2_policy_evict_first();
            for (long long u = ua; u < ub; ++u) {
                while (u >= pend) {
                    ++k;
                    pbeg = pend;
                    pend = T.panel_u0[k + 1];
                }
                mbar_wait(&empty_bar[stage], phase ^ 1u);
                const int b = (int)(u - pbeg);
                const int c0 = k * MV_THREADS;
                const int nf4 = min(MV_THREADS, n4 - c0);
                const unsigned int row_bytes = (unsigned int)nf4 * 16u;
                mbar_expect_tx(&full_bar[stage], row_bytes * R);
#pragma unroll
                for (int rr = 0; rr < R; ++rr) {
                    int r = b * R + rr;                 // local row
                    r = r < nloc ? r : nloc - 1;
                    if (T.l2_hint) bulk_g2s_hint(tiles + ((size_t)stage * R + rr) * MV_THREADS, A + (long long)r * ld + 4LL * c0, row_bytes, &full_bar[stage], pol);
                    else bulk_g2s(tiles + ((size_t)stage * R + rr) * MV_THREADS, A + (long long)r * ld + 4LL * c0, row_bytes, &full_bar[stage]);
                }
                if (++stage == STAGES) {
                    stage = 0;
                    phase ^= 1u;
                }
            }
        }
    } else if (MODE != SYM_BEGIN) {
    // ---------------- consumers: group g takes units u0+g, u0+g+NG, ... ----------------
    const int g = warp / WPG, wig = warp % WPG, tig = tid - g * TG;
    const double2* z2 = reinterpret_cast<const double2*>(z);
    const long long rp_ld = (long long)n4 * 4;
    int k = kfirst;
    long long pbeg = single_panel ? 0 : T.panel_u0[k];
    long long pend = single_panel ? NOEND : T.panel_u0[k + 1];
    double cacc[NQ][4], zc[NQ][4];
    int pfull = T.panel_full[k];
    auto load_zc = [&](int kk) {
        pfull = T.panel_full[kk];
#pragma unroll
        for (int q = 0; q < NQ; ++q) {
            const int c4 = kk * MV_THREADS + q * TG + tig;
            double2 za = make_double2(0.0, 0.0), zb = za;
            if (c4 < n4) {
                if (ZNC) { za = __ldg(z2 + 2 * c4); zb = __ldg(z2 + 2 * c4 + 1); }
                else { za = z2[2 * c4]; zb = z2[2 * c4 + 1]; }
            }
            zc[q][0] = za.x; zc[q][1] = za.y; zc[q][2] = zb.x; zc[q][3] = zb.y;
        }
    };
    // Column partials leave the CTA as ONE vector per (share or chunk, panel): the groups g > 0 hand theirs to group 0 through shared
    // memory (every group flushes the same panels in the same order, so the two consumer-wide barriers pair up), group 0 adds them
    // in group order and stores -- half the partials for the next launch's fold to read.
    auto flush = [&](int kk) {
        if (g > 0) {
#pragma unroll
            for (int q = 0; q < NQ; ++q) {
                double2* st = reinterpret_cast<double2*>(cred + (size_t)(g - 1) * 1024) + 2 * (q * TG + tig);
                st[0] = make_double2(cacc[q][0], cacc[q][1]);
                st[1] = make_double2(cacc[q][2], cacc[q][3]);
            }
        }
        asm volatile("bar.sync 1, %0;" ::"n"(SYM_CW * 32) : "memory");
        if (g == 0) {
#pragma unroll
            for (int q = 0; q < NQ; ++q) {
                double a0 = cacc[q][0], a1 = cacc[q][1], a2 = cacc[q][2], a3 = cacc[q][3];
#pragma unroll
                for (int gg = 1; gg < NG; ++gg) {
                    const double2* ld2 = reinterpret_cast<const double2*>(cred + (size_t)(gg - 1) * 1024) + 2 * (q * TG + tig);
                    const double2 x = ld2[0], y = ld2[1];
                    a0 += x.x; a1 += x.y; a2 += y.x; a3 += y.y;
                }
                double2* dst = reinterpret_cast<double2*>(colslots + ((size_t)(kk - kfirst) * NG) * 1024) + 2 * (q * TG + tig);
                dst[0] = make_double2(a0, a1);
                dst[1] = make_double2(a2, a3);
            }
        }
        asm volatile("bar.sync 1, %0;" ::"n"(SYM_CW * 32) : "memory");
#pragma unroll
        for (int q = 0; q < NQ; ++q) cacc[q][0] = cacc[q][1] = cacc[q][2] = cacc[q][3] = 0.0;
    };
#pragma unroll
    for (int q = 0; q < NQ; ++q) cacc[q][0] = cacc[q][1] = cacc[q][2] = cacc[q][3] = 0.0;
    load_zc(k);
    for (long long u = u0 + g; u < u1; u += NG) {
        while (u >= pend) {
            flush(k);
            ++k;
            pbeg = pend;
            pend = T.panel_u0[k + 1];
            load_zc(k);
        }
        const int it = (int)(u - u0);
        const int stage = it % STAGES;
        const unsigned int phase = (unsigned int)(it / STAGES) & 1u;
        const int l0 = (int)(u - pbeg) * R;          // local row of the band
        const int i0 = row0 + l0;                    // global row (row0 is a multiple of 4)
        // z of the band's rows (broadcast loads); rows past n read the zero padding of z
        double2 zr01, zr23;
        if (ZNC) { zr01 = __ldg(z2 + (i0 >> 1)); zr23 = __ldg(z2 + (i0 >> 1) + 1); }
        else { zr01 = z2[i0 >> 1]; zr23 = z2[(i0 >> 1) + 1]; }
        const double zr[4] = {zr01.x, zr01.y, zr23.x, zr23.y};
        mbar_wait(&full_bar[stage], phase);
        double rp[R] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
        for (int q = 0; q < NQ; ++q) {
            const int cl = q * TG + tig;                    // float4 column inside the panel
            const int c4 = k * MV_THREADS + cl;
            // j0 and i0 are multiples of 4: a thread's 4x4 block is entirely above the diagonal (delta > 0), entirely
            // below it (delta < 0: nothing to do) or the diagonal block itself (delta == 0, one thread per unit)
            // (an off-diagonal block assigned to this rank -- pfull -- counts every element for both parts)
            const int delta = pfull ? 4 : 4 * c4 - i0;
            if (c4 < n4 && delta >= 0) {
                float4 a[R];
#pragma unroll
                for (int rr = 0; rr < R; ++rr) a[rr] = tiles[((size_t)stage * R + rr) * MV_THREADS + cl];
                if (delta > 0) {
#pragma unroll
                    for (int rr = 0; rr < R; ++rr) {
                        const double x0 = (double)a[rr].x, x1 = (double)a[rr].y, x2 = (double)a[rr].z, x3 = (double)a[rr].w;
                        rp[rr] = fma(x0, zc[q][0], rp[rr]); rp[rr] = fma(x1, zc[q][1], rp[rr]);
                        rp[rr] = fma(x2, zc[q][2], rp[rr]); rp[rr] = fma(x3, zc[q][3], rp[rr]);
                        cacc[q][0] = fma(x0, zr[rr], cacc[q][0]); cacc[q][1] = fma(x1, zr[rr], cacc[q][1]);
                        cacc[q][2] = fma(x2, zr[rr], cacc[q][2]); cacc[q][3] = fma(x3, zr[rr], cacc[q][3]);
                    }
                } else {
                    // diagonal 4x4 block: row part takes j >= i, column part j > i
#pragma unroll
                    for (int rr = 0; rr < R; ++rr) {
                        const float av[4] = {a[rr].x, a[rr].y, a[rr].z, a[rr].w};
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            const double x = (double)av[e];
                            if (e >= rr) rp[rr] = fma(x, zc[q][e], rp[rr]);
                            if (e > rr) cacc[q][e] = fma(x, zr[rr], cacc[q][e]);
                        }
                    }
                }
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty_bar[stage]);
        // butterfly: 4 row partials x 32 lanes -> lanes {0,8,16,24} hold the warp sums of rows {0,1,2,3}
        {
            const bool hi16 = lane & 16, hi8 = lane & 8;
            double k0 = hi16 ? rp[2] : rp[0], k1 = hi16 ? rp[3] : rp[1];
            const double s0 = hi16 ? rp[0] : rp[2], s1 = hi16 ? rp[1] : rp[3];
            k0 += __shfl_xor_sync(0xffffffffu, s0, 16);
            k1 += __shfl_xor_sync(0xffffffffu, s1, 16);
            double y = hi8 ? k1 : k0;
            const double sy = hi8 ? k0 : k1;
            y += __shfl_xor_sync(0xffffffffu, sy, 8);
            y += __shfl_xor_sync(0xffffffffu, y, 4);
            y += __shfl_xor_sync(0xffffffffu, y, 2);
            y += __shfl_xor_sync(0xffffffffu, y, 1);
            // join the group's WPG warps through shared memory (double-buffered by unit parity: one named barrier per
            // unit), fixed order -> one partial per (panel, group, row)
            const int par = (it / NG) & 1;
            if ((lane & 7) == 0) gred[((par * NG + g) * WPG + wig) * R + (lane >> 3)] = y;
            asm volatile("bar.sync %0, %1;" ::"r"(2 + g), "n"(TG) : "memory");
            if (wig == 0 && lane < R) {
                double ssum = 0.0;
#pragma unroll
                for (int w = 0; w < WPG; ++w) ssum += gred[((par * NG + g) * WPG + w) * R + lane];
                // one partial per (panel, row): a band is walked by exactly one group, once per pass
                if (l0 + lane < nloc) rowpart[(long long)k * rp_ld + i0 + lane] = ssum;
            }
        }
    }
    // every group flushes every panel the CTA's share touches (groups that saw no unit of a panel flush zeros)
    while (true) {
        flush(k);
        if (k >= klast) break;
        ++k;
    }
    }   // consumers
    }   // u0 < u1
    if (MODE != SYM_BEGIN) __syncthreads();
}

// One CTA's whole job in a symmetric pass: its static share, then chunks of the dynamic pool until the pool is empty.
template <int NG, int STAGES, bool ZNC, bool BEGUN = false>
__device__ __forceinline__ void matvec_sym_core(const float* __restrict__ A, long long ld, int n4, const double* z, double* rowpart,
                                                double* colpart, const SymTab& T, float4* tiles, uint64_t* full_bar, uint64_t* empty_bar,
                                                double* gred, double* cred) {
    __shared__ int s_chunk;
    const long long c = blockIdx.x;
    const long long t_start = (T.dbg_cycles != nullptr && threadIdx.x == 0) ? clock64() : 0;
    matvec_sym_range<NG, STAGES, ZNC, BEGUN ? SYM_CONT : SYM_ALL>(A, ld, n4, z, rowpart, colpart + (size_t)c * T.nslots * NG * 1024, T, tiles, full_bar,
                                                                  empty_bar, gred, cred, T.cta_u0[c], T.cta_u0[c + 1], T.cta_kf[c], T.cta_kl[c], false);
    if (BEGUN && T.dbg_stamps != nullptr && threadIdx.x == 0 && T.dbg_stamps[c * 16 + 8] != 0) T.dbg_stamps[c * 16 + 5] = (long long)globaltimer_ns();
    int nchunks = 0;
    while (T.ndyn > 0) {
        if (threadIdx.x == 0) s_chunk = (int)atomicAdd(T.dyn_counter, 1u);
        __syncthreads();
        if (s_chunk >= T.ndyn) break;
        const int ch = T.dyn_order[s_chunk];
        const int k = T.dyn_k[ch], b0 = T.dyn_b0[ch], cnt = T.dyn_cnt[ch];
        // (the range function's leading barrier also protects s_chunk against the next fetch)
        matvec_sym_range<NG, STAGES, ZNC>(A, ld, n4, z, rowpart, T.colpart_dyn + (size_t)ch * NG * 1024, T, tiles, full_bar, empty_bar, gred, cred,
                                          b0, b0 + cnt, k, k, true);
        ++nchunks;
    }
    if (T.dbg_cycles != nullptr && threadIdx.x == 0) T.dbg_cycles[c] = clock64() - t_start;
    if (BEGUN && T.dbg_stamps != nullptr && threadIdx.x == 0 && T.dbg_stamps[c * 16 + 8] != 0) {
        T.dbg_stamps[c * 16 + 6] = (long long)globaltimer_ns();
        T.dbg_stamps[c * 16 + 7] = nchunks;
    }
}

template <int STAGES, int OCC>
__global__ void __launch_bounds__(SYM_THREADS, OCC)
k_matvec_sym_tma(const float* __restrict__ A, long long ld, int n4, const double* __restrict__ z, double* rowpart,
                 double* colpart, SymTab T, const int* __restrict__ done) {
    if (done != nullptr && *done) return;
    extern __shared__ __align__(128) unsigned char tma_smem[];
    __shared__ __align__(8) uint64_t full_bar[STAGES];
    __shared__ __align__(8) uint64_t empty_bar[STAGES];
    __shared__ double gred[2 * SYM_CW * SYM_R];
    __shared__ __align__(16) double cred[(SYM_NG - 1) * 1024];       // column partials of the groups g > 0 on their way to group 0
    matvec_sym_core<SYM_NG, STAGES, true>(A, ld, n4, z, rowpart, colpart, T, reinterpret_cast<float4*>(tma_smem), full_bar, empty_bar, gred, cred);
}

// This rank's contribution to t_i, computed by 8 cooperating lanes: lane j sums the row partials of consumer group
// j % NG over the panels ki + j / NG, + 8/NG, ... (only for the rank's own rows; a (panel, band) is written by ONE
// group, the other group's entry keeps its initial zero, so adding both is exact) and every 8th of the CTAs that walked
// panel ki for the column partials; a fixed 3-step shuffle tree joins them.  Must be called by all 8 lanes of an aligned
// group; every lane returns the sum.  On one rank this is t_i; sharded, the ranks' sums are added by an allreduce.
__device__ __forceinline__ double sym_fold_group8(int i, bool valid, int n4, const double* rowpart, const double* colpart, SymTab T) {
    const int j = threadIdx.x & 7;
    const int nch = (n4 + MV_THREADS - 1) / MV_THREADS;
    const long long rp_ld = (long long)n4 * 4;
    double s = 0.0;
    if (valid) {
        const int ki = i >> 10;
        if (i >= T.row0 && i < T.row0 + T.nloc) {
            // entries of panels (or bands) this rank never walks keep their initial zero, so the loop may cover them
            const int kstart = T.kfold_lo < 0 ? ki : T.kfold_lo;
#pragma unroll 4
            for (int k = kstart + j; k < nch; k += 8) s += __ldcg(rowpart + (long long)k * rp_ld + i);
        }
        const int clo = T.panel_cta[ki], chi = T.panel_cta_hi[ki];
#pragma unroll 2
        for (int c = clo + j; c <= chi; c += 8) {
            const double* cp = colpart + ((size_t)c * T.nslots + (ki - T.cta_kf[c])) * SYM_NG * 1024 + (i & 1023);
            s += __ldcg(cp);            // the groups' partials were combined at the flush: slot 0 holds the CTA's sum
        }
        if (T.ndyn > 0) {
            // column partials of the dynamic chunks of panel ki, in chunk order; two independent chains (fixed order) keep
            // more loads in flight
            const int d0 = T.panel_dyn0[ki], d1 = T.panel_dyn0[ki + 1];
            double s2 = 0.0;
            int ch = d0 + j;
#pragma unroll 2
            for (; ch + 8 < d1; ch += 16) {
                const double* cp = T.colpart_dyn + (size_t)ch * SYM_NG * 1024 + (i & 1023);
                s += __ldcg(cp);
                s2 += __ldcg(cp + (size_t)8 * SYM_NG * 1024);
            }
            if (ch < d1) s += __ldcg(T.colpart_dyn + (size_t)ch * SYM_NG * 1024 + (i & 1023));
            s += s2;
        }
    }
    s += __shfl_xor_sync(0xffffffffu, s, 1);
    s += __shfl_xor_sync(0xffffffffu, s, 2);
    s += __shfl_xor_sync(0xffffffffu, s, 4);
    return s;
}

// The same sum for the deferred-scalar pass, organised for memory-level parallelism instead of lanes per bin: LPB (2) adjacent
// lanes share a bin, so a warp covers 32 / LPB CONSECUTIVE bins and every load instruction of the warp reads whole 32-byte sectors
// of one partial array; a lane takes every LPB-th source of the flattened source lists (row partials (panel, group); column
// partials (CTA, group) of the static shares; column partials (chunk, group) of the dynamic pool), loads them in batches of
// FOLD_UN independent loads and adds them in a fixed order.  One round of the grid covers 16 x 2664 bins, so the vector phase
// of a pass is a handful of L2 round trips instead of three latency-bound rounds.  Every lane of the group returns the sum.
constexpr int FOLD_UN = 16;
template <int FOLD_LPB>
__device__ __forceinline__ double sym_fold_mlp(int i, bool valid, int n4, const double* rowpart, const double* colpart, const SymTab& T,
                                               bool* has_sources = nullptr) {
    const int j = threadIdx.x & (FOLD_LPB - 1);
    const int nch = (n4 + MV_THREADS - 1) / MV_THREADS;
    const long long rp_ld = (long long)n4 * 4;
    double s = 0.0;
    if (valid) {
        const int ki = i >> 10, col = i & 1023;
        // the three source lists, flattened into one index range so that every batch of loads is full
        int n_row = 0;
        const double* rbase = rowpart;
        const bool listed = T.n_fold_k > 0;
        if (i >= T.row0 && i < T.row0 + T.nloc) {
            const int kstart = T.kfold_lo < 0 ? ki : T.kfold_lo;
            n_row = listed ? T.n_fold_k : nch - kstart;
            rbase = rowpart + (long long)(listed ? 0 : kstart) * rp_ld + i;
        }
        const int clo = T.panel_cta[ki], chi = T.panel_cta_hi[ki];
        const int n_col = chi >= clo ? (chi - clo + 1) : 0;            // one combined partial per CTA (slot of group 0)
        // only the first CTA of a panel can have started its share in an earlier panel (shares are contiguous and ordered)
        const int slot_lo = n_col > 0 ? ki - T.cta_kf[clo] : 0;
        const double* cbase = colpart + (size_t)clo * T.nslots * SYM_NG * 1024 + col;
        int n_dyn = 0;
        const double* dbase = T.colpart_dyn;
        if (T.ndyn > 0) {
            const int d0 = T.panel_dyn0[ki];
            n_dyn = T.panel_dyn0[ki + 1] - d0;
            dbase = T.colpart_dyn + (size_t)d0 * SYM_NG * 1024 + col;
        }
        const int n_rc = n_row + n_col, total = n_rc + n_dyn;
        if (has_sources != nullptr) *has_sources = total > 0;        // a static property of (rank, bin): the same in every pass
        const size_t cstride = (size_t)T.nslots * SYM_NG * 1024;
        for (int q = j; q < total; q += FOLD_LPB * FOLD_UN) {
            double v[FOLD_UN];
#pragma unroll
            for (int u = 0; u < FOLD_UN; ++u) {
                const int idx = q + u * FOLD_LPB;
                const double* p = nullptr;
                if (idx < n_row) {
                    p = rbase + (long long)(listed ? T.fold_k[idx] : idx) * rp_ld;
                } else if (idx < n_rc) {
                    const int cc = idx - n_row;
                    p = cbase + (size_t)cc * cstride + (size_t)(cc == 0 ? slot_lo : 0) * SYM_NG * 1024;
                } else if (idx < total) {
                    p = dbase + (size_t)(idx - n_rc) * SYM_NG * 1024;
                }
                v[u] = p != nullptr ? __ldcg(p) : 0.0;
            }
#pragma unroll
            for (int u = 0; u < FOLD_UN; ++u) s += v[u];
        }
    }
#pragma unroll
    for (int o = 1; o < FOLD_LPB; o <<= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    return s;
}

__global__ void __launch_bounds__(256)
k_sym_fold(int n, int n4, const double* rowpart, const double* colpart, SymTab T, double* __restrict__ t, const int* __restrict__ done) {
    if (done != nullptr && *done) return;
    const int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 3;
    const double ti = sym_fold_group8(i, i < n, n4, rowpart, colpart, T);
    if (i < n && (threadIdx.x & 7) == 0) t[i] = ti;
}

// Exact symmetry test: number of (i, j), i < j, with A_ij != A_ji (bitwise on the float values; +0 == -0).
// One read of the map (32 x 32 tiles, the mirror tile transposed through shared memory).
__global__ void __launch_bounds__(256)
k_count_asym(const float* __restrict__ A, long long ld, int n, unsigned long long* count) {
    __shared__ float tile[32][33];
    const int nt = (n + 31) / 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;      // 32 x 8
    unsigned int bad = 0;
    for (long long p = blockIdx.x; p < (long long)nt * nt; p += gridDim.x) {
        const int bi = (int)(p / nt), bj = (int)(p % nt);
        if (bj < bi) continue;
        __syncthreads();
        for (int r = ty; r < 32; r += 8) {            // mirror tile (bj, bi)
            const int i = bj * 32 + r, j = bi * 32 + tx;
            tile[r][tx] = (i < n && j < n) ? A[(long long)i * ld + j] : 0.f;
        }
        __syncthreads();
        for (int r = ty; r < 32; r += 8) {
            const int i = bi * 32 + r, j = bj * 32 + tx;
            if (i < n && j < n && i < j) bad += (A[(long long)i * ld + j] != tile[tx][r]);
        }
    }
    bad = __reduce_add_sync(0xffffffffu, bad);
    if (tx == 0 && bad) atomicAdd(count, (unsigned long long)bad);
}

// Symmetry test for a ROW-SHARDED map, where no rank sees a block and its mirror image: every rank adds, over its own rows, a 64-bit
// hash of (min(i,j), max(i,j), value bits) for the entries above the diagonal into one sum and for the entries below it into
// another (+0 == -0: zeros are skipped).  The map is symmetric iff the multiset of (pair, value) above the diagonal equals the one
// below, so the two sums -- added over the ranks, modulo 2^64 -- agree (and differ with probability 1 - 2^-64 per mismatching
// entry otherwise).  One read of the block; replaces the caller's promise ("assume_symmetric") of round 1.
__global__ void __launch_bounds__(MV_THREADS)
k_sym_hash(const float* __restrict__ A, long long ld, int nloc, int n, int row0, unsigned long long* out /* [2]: upper, lower */) {
    unsigned long long up = 0, lo = 0;
    for (int r = blockIdx.x; r < nloc; r += gridDim.x) {
        const int i = row0 + r;
        const float4* rp = reinterpret_cast<const float4*>(A + (long long)r * ld);
        const int n4 = (n + 3) / 4;
        for (int c4 = threadIdx.x; c4 < n4; c4 += MV_THREADS) {
            const float4 a = ld_stream_f4(rp + c4);
            const float e[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int j = 4 * c4 + k;
                const unsigned int bits = __float_as_uint(e[k]);
                if (j < n && j != i && (bits & 0x7fffffffu)) {
                    const uint64_t a0 = (uint64_t)(i < j ? i : j), a1 = (uint64_t)(i < j ? j : i);
                    const uint64_t h = mix64(mix64((a0 << 32) | a1) ^ (uint64_t)bits);
                    if (j > i) up += h; else lo += h;
                }
            }
        }
    }
    up = (unsigned long long)warp_sum_ll((long long)up);
    lo = (unsigned long long)warp_sum_ll((long long)lo);
    if ((threadIdx.x & 31) == 0) {
        atomicAdd(out, up);
        atomicAdd(out + 1, lo);
    }
}

// ---------------------------------------------------------------------------------------------------
// K1 row statistics: float64 row sum + int32 count of zeros over the n real columns.
// Replaces the loop of lib/ccmapHelpers.pyx:190-201.  Algorithmic bytes: 4 * nloc * n.
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(MV_THREADS)
k_row_stats(const float* __restrict__ A, long long ld, int nloc, int n, double* __restrict__ rowsum,
            int* __restrict__ zero_count, int* __restrict__ neg_count = nullptr) {
    __shared__ double sm[32];
    const int n4 = (int)(ld / 4);
    const int pad = (int)ld - n;
    for (int r = blockIdx.x; r < nloc; r += gridDim.x) {
        const float4* rp = reinterpret_cast<const float4*>(A + (long long)r * ld);
        double s = 0.0;
        int zc = 0, ng = 0;
        for (int c4 = threadIdx.x; c4 < n4; c4 += 4 * MV_THREADS) {
            float4 a[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int cc = c4 + k * MV_THREADS;
                a[k] = cc < n4 ? ld_stream_f4(rp + cc) : make_float4(1.f, 1.f, 1.f, 1.f);
            }
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int cc = c4 + k * MV_THREADS;
                if (cc < n4) {
                    s += ((double)a[k].x + (double)a[k].y) + ((double)a[k].z + (double)a[k].w);
                    zc += (a[k].x == 0.f) + (a[k].y == 0.f) + (a[k].z == 0.f) + (a[k].w == 0.f);
                    ng += (a[k].x < 0.f) + (a[k].y < 0.f) + (a[k].z < 0.f) + (a[k].w < 0.f);
                }
            }
        }
        const double S = block_reduce<RED_SUM>(s, sm);
        const double Z = block_reduce<RED_SUM>((double)zc, sm);   // exact: counts < 2^53
        // rows with negative entries: the reference's emptiness test is a FLOAT32 sum (lib/ccmapHelpers.pyx:192), which can cancel to
        // exactly 0 where the float64 sum does not; the host re-checks just those rows in NumPy's float32 arithmetic
        const int any_neg = __syncthreads_or(ng);
        if (neg_count != nullptr && threadIdx.x == 0) neg_count[r] = any_neg ? 1 : 0;
        if (threadIdx.x == 0) {
            rowsum[r] = S;
            // raw count of zeros among the n real columns; the host applies lib/ccmapHelpers.pyx:192-195
            // (a row whose sum is 0.0 records a zero count of 0)
            zero_count[r] = (int)Z - pad;
        }
    }
}

// ---------------------------------------------------------------------------------------------------
// elementwise passes: same (band, chunk) decomposition, column-vector entries held in registers
// ---------------------------------------------------------------------------------------------------
struct MinMax {
    unsigned int min_nz;   // f2ord-encoded min over non-zero entries (init 0xffffffff)
    unsigned int max_all;  // f2ord-encoded max (init 0)
};

// per-thread tracking stays in the float domain (one FMNMX each); the order-preserving encoding is applied
// once per thread at commit.  mn starts at +inf (no non-zero seen), mx at -inf (nothing seen).
#define GCX_PINF __int_as_float(0x7f800000)
#define GCX_NINF __int_as_float(0xff800000)
__device__ __forceinline__ void mm_update(float v, float& mn, float& mx) {
    if (v != 0.f) mn = fminf(mn, v);
    mx = fmaxf(mx, v);
}

__device__ __forceinline__ void mm_commit(float fmn, float fmx, MinMax* out) {
    unsigned int mn = fmn == GCX_PINF ? 0xffffffffu : f2ord(fmn);
    unsigned int mx = fmx == GCX_NINF ? 0u : f2ord(fmx);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        mn = min(mn, __shfl_xor_sync(0xffffffffu, mn, o));
        mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    }
    if ((threadIdx.x & 31) == 0) {
        if (mn != 0xffffffffu) atomicMin(&out->min_nz, mn);
        atomicMax(&out->max_all, mx);
    }
}

// K4: out_ij = f32((s_lo * A_ij) * s_hi) for kept i,j (lo/hi by global index), else 0 or pass-through.
// Replaces GenNormMatrixRAM (lib/normalizeKnightRuiz.pyx:207-227) and the net effect of
// lib/normalizeCore.pyx:48-53 + the write-back of lib/normalizer.py:601-613.
// Algorithmic bytes: 8 * nloc * n.  min/max over the kept x kept block only.
template <int R, int OCC>
__global__ void __launch_bounds__(MV_THREADS, OCC)
k_apply_sym(const float* __restrict__ A, float* __restrict__ Out, long long ld, int nloc, int n4, int row0,
            const double* __restrict__ s, const uint8_t* __restrict__ keep, int passthrough, MinMax* mm) {
    const int tid = threadIdx.x;
    const int nch = (n4 + MV_THREADS - 1) / MV_THREADS;
    const int nb = (nloc + R - 1) / R;
    const long long U = (long long)nb * nch, G = gridDim.x, c = blockIdx.x;
    const long long u0 = unit_begin(c, U, G), u1 = unit_begin(c + 1, U, G);
    float mn = GCX_PINF, mx = GCX_NINF;
    const double2* __restrict__ s2 = reinterpret_cast<const double2*>(s);
    const uchar4* __restrict__ k4 = reinterpret_cast<const uchar4*>(keep);
    for (long long u = u0; u < u1; ++u) {
        const int b = (int)(u / nch), k = (int)(u - (long long)b * nch);
        const int c4 = k * MV_THREADS + tid;
        if (c4 >= n4) continue;
        float4 a[R];
#pragma unroll
        for (int rr = 0; rr < R; ++rr) {
            const int r = b * R + rr;
            if (r < nloc) a[rr] = ld_stream_f4(reinterpret_cast<const float4*>(A + (long long)r * ld) + c4);
        }
        const double2 sa = __ldg(s2 + 2 * c4), sb = __ldg(s2 + 2 * c4 + 1);
        const uchar4 kc = __ldg(k4 + c4);
        const double sc[4] = {sa.x, sa.y, sb.x, sb.y};
        const unsigned char kcol[4] = {kc.x, kc.y, kc.z, kc.w};
        const int j0 = 4 * c4;
#pragma unroll
        for (int rr = 0; rr < R; ++rr) {
            const int r = b * R + rr;
            if (r >= nloc) break;
            const int i = row0 + r;
            const double si = __ldg(s + i);
            const bool ki = __ldg(keep + i) != 0;
            const float in[4] = {a[rr].x, a[rr].y, a[rr].z, a[rr].w};
            float o[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int j = j0 + e;
                if (ki && kcol[e]) {
                    if (in[e] != 0.f) {     // (lo*0)*hi = 0: zeros (most of a Hi-C map) skip the fp64 work
                        const double lo = j < i ? sc[e] : si, hi = j < i ? si : sc[e];
                        o[e] = (float)((lo * (double)in[e]) * hi);
                    } else {
                        o[e] = in[e];
                    }
                    mm_update(o[e], mn, mx);
                } else {
                    o[e] = passthrough ? in[e] : 0.f;
                }
            }
            st_stream_f4(reinterpret_cast<float4*>(Out + (long long)r * ld) + c4, make_float4(o[0], o[1], o[2], o[3]));
        }
    }
    mm_commit(mn, mx, mm);
}

// K5: vanilla coverage in the reference's fp32 arithmetic (lib/normalizeCore.pyx:72-86):
// out = A; kept row i: out = fl32(out / c_i); kept column j: out = fl32(out / c_j); optional sqrt.
// Algorithmic bytes: 8 * nloc * n (+ the row-sum pass of k_row_stats = 12 n^2 in total).
template <int R, int OCC>
__global__ void __launch_bounds__(MV_THREADS, OCC)
k_vc_apply(const float* __restrict__ A, float* __restrict__ Out, long long ld, int nloc, int n4, int n, int row0,
           const float* __restrict__ csum, const uint8_t* __restrict__ keep, int sqroot, MinMax* mm) {
    const int tid = threadIdx.x;
    const int nch = (n4 + MV_THREADS - 1) / MV_THREADS;
    const int nb = (nloc + R - 1) / R;
    const long long U = (long long)nb * nch, G = gridDim.x, c = blockIdx.x;
    const long long u0 = unit_begin(c, U, G), u1 = unit_begin(c + 1, U, G);
    float mn = GCX_PINF, mx = GCX_NINF;
    const float4* __restrict__ c4v = reinterpret_cast<const float4*>(csum);
    const uchar4* __restrict__ k4 = reinterpret_cast<const uchar4*>(keep);
    for (long long u = u0; u < u1; ++u) {
        const int b = (int)(u / nch), k = (int)(u - (long long)b * nch);
        const int c4 = k * MV_THREADS + tid;
        if (c4 >= n4) continue;
        float4 a[R];
#pragma unroll
        for (int rr = 0; rr < R; ++rr) {
            const int r = b * R + rr;
            if (r < nloc) a[rr] = ld_stream_f4(reinterpret_cast<const float4*>(A + (long long)r * ld) + c4);
        }
        const float4 cs = __ldg(c4v + c4);
        const uchar4 kc = __ldg(k4 + c4);
        const float cc[4] = {cs.x, cs.y, cs.z, cs.w};
        const unsigned char kcol[4] = {kc.x, kc.y, kc.z, kc.w};
        const int j0 = 4 * c4;
#pragma unroll
        for (int rr = 0; rr < R; ++rr) {
            const int r = b * R + rr;
            if (r >= nloc) break;
            const int i = row0 + r;
            const float ci = __ldg(csum + i);
            const bool ki = __ldg(keep + i) != 0;
            const float in[4] = {a[rr].x, a[rr].y, a[rr].z, a[rr].w};
            float o[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                float v = in[e];
                if (v != 0.f) {     // 0/c = 0 exactly; skipping keeps zeros (most of a Hi-C map) off fdiv's slow path
                    if (ki) v = __fdiv_rn(v, ci);
                    if (kcol[e]) v = __fdiv_rn(v, cc[e]);
                    if (sqroot) v = __fsqrt_rn(v);
                }
                o[e] = v;
                if (j0 + e < n) mm_update(v, mn, mx);
            }
            st_stream_f4(reinterpret_cast<float4*>(Out + (long long)r * ld) + c4, make_float4(o[0], o[1], o[2], o[3]));
        }
    }
    mm_commit(mn, mx, mm);
}

// .hic import with a precomputed norm vector pair (second half): every stored count becomes
// count / (c1[bin_x] * c2[bin_y]), +inf where that product is zero (lib/hic2gcmap.py:88-95, the vectors of
// lib/hicparser.py:358-412); the arithmetic is Python's float64, the map float32.  Bins without a record (zeros of the dense map)
// stay zero.  Algorithmic bytes: 8 * nloc * n.
template <int R, int OCC>
__global__ void __launch_bounds__(MV_THREADS, OCC)
k_apply_norm_vectors(const float* __restrict__ A, float* __restrict__ Out, long long ld, int nloc, int n4, int n, int row0,
                     const double* __restrict__ c1, const double* __restrict__ c2, MinMax* mm) {
    const int tid = threadIdx.x;
    const int nch = (n4 + MV_THREADS - 1) / MV_THREADS;
    const int nb = (nloc + R - 1) / R;
    const long long U = (long long)nb * nch, G = gridDim.x, c = blockIdx.x;
    const long long u0 = unit_begin(c, U, G), u1 = unit_begin(c + 1, U, G);
    float mn = GCX_PINF, mx = GCX_NINF;
    const double2* __restrict__ c22 = reinterpret_cast<const double2*>(c2);
    for (long long u = u0; u < u1; ++u) {
        const int b = (int)(u / nch), k = (int)(u - (long long)b * nch);
        const int c4 = k * MV_THREADS + tid;
        if (c4 >= n4) continue;
        float4 a[R];
#pragma unroll
        for (int rr = 0; rr < R; ++rr) {
            const int r = b * R + rr;
            if (r < nloc) a[rr] = ld_stream_f4(reinterpret_cast<const float4*>(A + (long long)r * ld) + c4);
        }
        const double2 ca = __ldg(c22 + 2 * c4), cb = __ldg(c22 + 2 * c4 + 1);
        const double cj[4] = {ca.x, ca.y, cb.x, cb.y};
        const int j0 = 4 * c4;
#pragma unroll
        for (int rr = 0; rr < R; ++rr) {
            const int r = b * R + rr;
            if (r >= nloc) break;
            const double ci = __ldg(c1 + row0 + r);
            const float in[4] = {a[rr].x, a[rr].y, a[rr].z, a[rr].w};
            float o[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                float v = in[e];
                if (v != 0.f && j0 + e < n) {
                    const double norm = ci * cj[e];                                   // lib/hic2gcmap.py:89
                    v = norm != 0.0 ? (float)((double)in[e] / norm) : GCX_PINF;       // :90-93
                }
                o[e] = v;
                if (j0 + e < n) mm_update(v, mn, mx);
            }
            st_stream_f4(reinterpret_cast<float4*>(Out + (long long)r * ld) + c4, make_float4(o[0], o[1], o[2], o[3]));
        }
    }
    mm_commit(mn, mx, mm);
}

// K4 + K5 fused: one read of the map, two results -- the IC (or KR) scaled map of k_apply_sym into OutS and the vanilla-coverage
// map of k_vc_apply into OutV, each with its own mask and its own min/max.  Same arithmetic per element as the two kernels
// (bit-identical outputs); algorithmic bytes 12 * nloc * n instead of 16 * nloc * n for the pair.
template <int R, int OCC>
__global__ void __launch_bounds__(MV_THREADS, OCC)
k_apply_sym_vc(const float* __restrict__ A, float* __restrict__ OutS, float* __restrict__ OutV, long long ld, int nloc, int n4, int n, int row0,
               const double* __restrict__ s, const uint8_t* __restrict__ keep_s, int passthrough, const float* __restrict__ csum,
               const uint8_t* __restrict__ keep_v, int sqroot, MinMax* mm_s, MinMax* mm_v) {
    const int tid = threadIdx.x;
    const int nch = (n4 + MV_THREADS - 1) / MV_THREADS;
    const int nb = (nloc + R - 1) / R;
    const long long U = (long long)nb * nch, G = gridDim.x, c = blockIdx.x;
    const long long u0 = unit_begin(c, U, G), u1 = unit_begin(c + 1, U, G);
    float mn_s = GCX_PINF, mx_s = GCX_NINF, mn_v = GCX_PINF, mx_v = GCX_NINF;
    const double2* __restrict__ s2 = reinterpret_cast<const double2*>(s);
    const uchar4* __restrict__ ks4 = reinterpret_cast<const uchar4*>(keep_s);
    const uchar4* __restrict__ kv4 = reinterpret_cast<const uchar4*>(keep_v);
    const float4* __restrict__ c4v = reinterpret_cast<const float4*>(csum);
    for (long long u = u0; u < u1; ++u) {
        const int b = (int)(u / nch), k = (int)(u - (long long)b * nch);
        const int c4 = k * MV_THREADS + tid;
        if (c4 >= n4) continue;
        float4 a[R];
#pragma unroll
        for (int rr = 0; rr < R; ++rr) {
            const int r = b * R + rr;
            if (r < nloc) a[rr] = ld_stream_f4(reinterpret_cast<const float4*>(A + (long long)r * ld) + c4);
        }
        const double2 sa = __ldg(s2 + 2 * c4), sb = __ldg(s2 + 2 * c4 + 1);
        const uchar4 ksc = __ldg(ks4 + c4), kvc = __ldg(kv4 + c4);
        const float4 cs = __ldg(c4v + c4);
        const double sc[4] = {sa.x, sa.y, sb.x, sb.y};
        const float cc[4] = {cs.x, cs.y, cs.z, cs.w};
        const unsigned char kscol[4] = {ksc.x, ksc.y, ksc.z, ksc.w}, kvcol[4] = {kvc.x, kvc.y, kvc.z, kvc.w};
        const int j0 = 4 * c4;
#pragma unroll
        for (int rr = 0; rr < R; ++rr) {
            const int r = b * R + rr;
            if (r >= nloc) break;
            const int i = row0 + r;
            const double si = __ldg(s + i);
            const float ci = __ldg(csum + i);
            const bool ksi = __ldg(keep_s + i) != 0, kvi = __ldg(keep_v + i) != 0;
            const float in[4] = {a[rr].x, a[rr].y, a[rr].z, a[rr].w};
            float os[4], ov[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int j = j0 + e;
                const float x = in[e];
                // ---- symmetric scale (k_apply_sym) ----
                if (ksi && kscol[e]) {
                    if (x != 0.f) {
                        const double lo = j < i ? sc[e] : si, hi = j < i ? si : sc[e];
                        os[e] = (float)((lo * (double)x) * hi);
                    } else {
                        os[e] = x;
                    }
                    mm_update(os[e], mn_s, mx_s);
                } else {
                    os[e] = passthrough ? x : 0.f;
                }
                // ---- vanilla coverage (k_vc_apply) ----
                float v = x;
                if (v != 0.f) {
                    if (kvi) v = __fdiv_rn(v, ci);
                    if (kvcol[e]) v = __fdiv_rn(v, cc[e]);
                    if (sqroot) v = __fsqrt_rn(v);
                }
                ov[e] = v;
                if (j < n) mm_update(v, mn_v, mx_v);
            }
            st_stream_f4(reinterpret_cast<float4*>(OutS + (long long)r * ld) + c4, make_float4(os[0], os[1], os[2], os[3]));
            st_stream_f4(reinterpret_cast<float4*>(OutV + (long long)r * ld) + c4, make_float4(ov[0], ov[1], ov[2], ov[3]));
        }
    }
    mm_commit(mn_s, mx_s, mm_s);
    mm_commit(mn_v, mx_v, mm_v);
}

// K7: distance-frequency apply (lib/normalizeCore.pyx:91-170): divide or subtract by avg[|i-j|].
// Algorithmic bytes: 8 * nloc * n.
//
// The reference divides in float64 and stores float32: out = f32(RN64(x / av)).  A float64 division is ~30 instructions with
// its special cases, and on a sparse map nearly every warp has a few non-zero lanes, so the whole warp pays it for every entry
// (the first version of the batched kernel ran 47 instructions per entry at 17 of 32 lanes active).  Here every entry takes the
// same branch-free path: with r = RN64(1 / av) tabulated per distance next to av,
//     q0 = x r,  e = fma(-q0, av, x),  q = fma(e, r, q0)
// is RN64(x / av) or its neighbour (|q - x/av| < 2^-104 |x/av| before the last rounding).  The two can only round to different
// float32 values when q sits within one float64 ulp of the midpoint between two floats; those entries (29 low mantissa bits within
// 0x10000000 +- 1: about one in 10^8), results below the float32 normal range and non-finite intermediates take the exact
// division.  The stored float is therefore the reference's, bit for bit (tests/test_gpu_parity.py::test_diag_apply_exactness).
__global__ void k_diag_recip(int n, const double* __restrict__ avg, double2* __restrict__ avr) {
    const int d = blockIdx.x * blockDim.x + threadIdx.x;
    if (d < n) {
        const double a = avg[d];
        avr[d] = make_double2(a, a != 0.0 ? 1.0 / a : 0.0);
    }
}

// out of line on purpose: inlined, the compiler hoists the division above the test and every entry pays it again
__device__ __noinline__ double diag_exact_quotient(double x, double av) { return x / av; }

// Four adjacent entries of row i (columns j0 .. j0 + 3): one test for the rare exact-division path instead of one per entry.
// CS: column stride between the four entries, 1 (a float4 of the row) or 32 (lane-strided, see k_diag_apply_lane_b); SUB: o - e
// instead of o / e; GUARD: some of the four columns may lie in the padding (j >= n).
template <int CS, bool SUB, bool GUARD>
__device__ __forceinline__ void diag_apply_quad(const float (&x)[4], float (&o)[4], int i, int j0, int n, const double2* __restrict__ avr,
                                                float& mn, float& mx) {
    double q[4];
    bool live[4];
    unsigned int risky = 0u;
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        const int j = j0 + e * CS;
        const int d = GUARD ? min(abs(i - j), n - 1) : abs(i - j);       // padding columns read a valid slot; they hold zeros
        const double2 ar = __ldg(avr + d);
        const double xd = (double)x[e];
        live[e] = x[e] != 0.f && ar.x != 0.0;                 // a zero entry stays zero in both modes (0/avg = 0; :142-144 forces 0 for o-e)
        if (!SUB) {
            const double q0 = xd * ar.y;
            const double r = fma(-q0, ar.x, xd);
            q[e] = fma(r, ar.y, q0);                                                   // :105-108, :114-117
            const unsigned int lo = (unsigned int)__double2loint(q[e]), ex = ((unsigned int)__double2hiint(q[e]) >> 20) & 0x7ffu;
            const bool rk = ((lo & 0x1fffffffu) - 0x0fffffffu) <= 2u || (ex - 899u) >= (0x7ffu - 899u);
            risky |= (live[e] && rk) ? (1u << e) : 0u;
        } else {
            const double t = xd - ar.x;                                                // :138
            q[e] = t < 0.0 ? 1.0 : t;                                                  // :139-141
        }
    }
    if (!SUB && risky) {
#pragma unroll
        for (int e = 0; e < 4; ++e)
            if (risky & (1u << e)) q[e] = diag_exact_quotient((double)x[e], __ldg(avr + min(abs(i - j0 - e * CS), n - 1)).x);
    }
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        const float v = live[e] ? (float)q[e] : x[e];
        o[e] = v;
        const bool in = !GUARD || j0 + e * CS < n;
        mn = (in && v != 0.f) ? fminf(mn, v) : mn;
        mx = in ? fmaxf(mx, v) : mx;
    }
}

template <int R, int OCC>
__global__ void __launch_bounds__(MV_THREADS, OCC)
k_diag_apply(const float* __restrict__ A, float* __restrict__ Out, long long ld, int nloc, int n4, int n, int row0,
             const double2* __restrict__ avr, int subtract, MinMax* mm) {
    const int tid = threadIdx.x;
    const int nch = (n4 + MV_THREADS - 1) / MV_THREADS;
    const int nb = (nloc + R - 1) / R;
    const long long U = (long long)nb * nch, G = gridDim.x, c = blockIdx.x;
    const long long u0 = unit_begin(c, U, G), u1 = unit_begin(c + 1, U, G);
    float mn = GCX_PINF, mx = GCX_NINF;
    for (long long u = u0; u < u1; ++u) {
        const int b = (int)(u / nch), k = (int)(u - (long long)b * nch);
        const int c4 = k * MV_THREADS + tid;
        if (c4 >= n4) continue;
        float4 a[R];
#pragma unroll
        for (int rr = 0; rr < R; ++rr) {
            const int r = b * R + rr;
            if (r < nloc) a[rr] = ld_stream_f4(reinterpret_cast<const float4*>(A + (long long)r * ld) + c4);
        }
        const int j0 = 4 * c4;
#pragma unroll
        for (int rr = 0; rr < R; ++rr) {
            const int r = b * R + rr;
            if (r >= nloc) break;
            const int i = row0 + r;
            const float in[4] = {a[rr].x, a[rr].y, a[rr].z, a[rr].w};
            float o[4];
            if (subtract) diag_apply_quad<1, true, true>(in, o, i, j0, n, avr, mn, mx);
            else diag_apply_quad<1, false, true>(in, o, i, j0, n, avr, mn, mx);
            st_stream_f4(reinterpret_cast<float4*>(Out + (long long)r * ld) + c4, make_float4(o[0], o[1], o[2], o[3]));
        }
    }
    mm_commit(mn, mx, mm);
}

// ---------------------------------------------------------------------------------------------------
// K6: per-diagonal statistics of the non-zero entries (lib/cmstats.py:569-605).
// Only the lower triangle is read (the reference gathers rows[d:], cols[:-d], lib/util.py:95-96).
// Traversal: CTA (x = chunk of 256 diagonals, y = band of DIAG_RB rows); thread t owns diagonal
// d = 256*x + t and walks the band's rows i, reading A[i][i-d] -- per row a warp reads 32 consecutive
// floats (coalesced, descending).  Algorithmic bytes: 2 * n^2 per traversal (half the map).
// ---------------------------------------------------------------------------------------------------
constexpr int DIAG_RB = 64;
constexpr int DIAG_BATCH = 16;     // loads in flight per thread in the histogram traversal

// mean: per-(band, diagonal) partial sum/count, then a fixed-order reduction over bands (deterministic)
__device__ __forceinline__ void diag_mean_partial_body(const float* __restrict__ A, long long ld, int nloc, int n, int row0, double* __restrict__ psum,
                                                       int* __restrict__ pcnt, int bx, int by) {
    const int band = by, r0 = band * DIAG_RB, r1 = min(r0 + DIAG_RB, nloc);
    const int imax = row0 + r1 - 1;
    if ((int)(bx * MV_THREADS) > imax) return;
    const int d = bx * MV_THREADS + threadIdx.x;
    if (d > imax || d >= n) return;
    double s = 0.0;
    int c = 0;
    const int rs = max(r0, d - row0);
#pragma unroll 8
    for (int r = rs; r < r1; ++r) {
        const float v = __ldg(A + (long long)r * ld + (row0 + r - d));
        if (v != 0.f) {
            s += (double)v;
            ++c;
        }
    }
    psum[(long long)band * n + d] = s;
    pcnt[(long long)band * n + d] = c;
}

__global__ void __launch_bounds__(MV_THREADS)
k_diag_mean_partial(const float* __restrict__ A, long long ld, int nloc, int n, int row0, double* __restrict__ psum,
                    int* __restrict__ pcnt) {
    diag_mean_partial_body(A, ld, nloc, n, row0, psum, pcnt, blockIdx.x, blockIdx.y);
}

// 64 diagonals x 4 band slices per CTA; slice s adds bands b0+s, b0+s+4, ... in order, the four slice sums are
// combined in slice order -> the same result on every launch and for every grid shape
constexpr int DIAG_RED_SL = 4;
__device__ __forceinline__ void diag_mean_reduce_body(int nbands, int n, int row0, const double* __restrict__ psum, const int* __restrict__ pcnt,
                                                      double* __restrict__ sum, double* __restrict__ cnt, int bx) {
    __shared__ double ss[DIAG_RED_SL][64];
    __shared__ double sc[DIAG_RED_SL][64];
    const int dl = threadIdx.x & 63, sl = threadIdx.x >> 6;
    const int d = bx * 64 + dl;
    double s = 0.0, c = 0.0;
    if (d < n) {
        const int b0 = max(0, d - row0) / DIAG_RB;
#pragma unroll 8
        for (int b = b0 + sl; b < nbands; b += DIAG_RED_SL) {
            s += psum[(long long)b * n + d];
            c += (double)pcnt[(long long)b * n + d];
        }
    }
    ss[sl][dl] = s;
    sc[sl][dl] = c;
    __syncthreads();
    if (sl == 0 && d < n) {
#pragma unroll
        for (int k = 1; k < DIAG_RED_SL; ++k) { s += ss[k][dl]; c += sc[k][dl]; }
        sum[d] = s;
        cnt[d] = c;
    }
}

__global__ void __launch_bounds__(64 * DIAG_RED_SL)
k_diag_mean_reduce(int nbands, int n, int row0, const double* __restrict__ psum, const int* __restrict__ pcnt,
                   double* __restrict__ sum, double* __restrict__ cnt) {
    diag_mean_reduce_body(nbands, n, row0, psum, pcnt, sum, cnt, blockIdx.x);
}

__global__ void k_diag_mean_final(int n, const double* __restrict__ sum, const double* __restrict__ cnt, double* __restrict__ avg) {
    const int d = blockIdx.x * blockDim.x + threadIdx.x;
    if (d < n) avg[d] = cnt[d] > 0.0 ? sum[d] / cnt[d] : 0.0;      // lib/cmstats.py:598, :602-603
}

// median: MSB-first radix select on the order-preserving uint32 encoding of the floats, one histogram row per diagonal (global
// atomics; lanes of a warp hit 32 different diagonals).  Digits of 9 + 9 + 7 + 7 bits: the first two passes fix the sign, the
// exponent and nine mantissa bits -- every integer count up to 1023 exactly.  While pass 1 runs, the low 14 bits of the candidates
// (the entries inside the bucket pass 0 selected) are OR-ed per diagonal; when they are all zero the second selection is final and
// the diagonal leaves the search: on count maps the median costs TWO reads of the triangle, and passes 2 / 3 return at once when
// no diagonal is pending.  np.median of an even count needs the next larger value as well: every pass also tracks (atomicMin) the
// entries of the bucket next to the selected one at the level above, the last selection adds the next occupied bin of its own
// histogram -- no extra pass over the map.  A thread adds its entries of a band into a 4-entry (bin, count) cache in registers
// before it touches the histogram (count data repeats a handful of values along a diagonal).
constexpr int DIAG_NBIN = 512;                 // histogram row (passes 2, 3 use the first 128 bins)
constexpr int DIAG_HB = 256;                   // rows per CTA of the histogram passes: long enough that filling the slots is rare
constexpr int DIAG_SEL_ARRAYS = 8;
__host__ __device__ __forceinline__ int diag_digit_bits(int pass) { return pass < 2 ? 9 : 7; }
__host__ __device__ __forceinline__ int diag_digit_shift(int pass) { return pass == 0 ? 23 : (pass == 1 ? 14 : (pass == 2 ? 7 : 0)); }
struct DiagSel {
    unsigned int* hist;     // [n][DIAG_NBIN]
    unsigned int* prefix;   // [n] high bits fixed so far (right-aligned); the full key once the diagonal is done
    unsigned int* krem;     // [n] rank still to skip inside the current prefix group
    unsigned int* cnt;      // [n] number of non-zero entries
    unsigned int* need;     // [n] 1 if the upper middle value is the next distinct key
    unsigned int* next;     // [n] min key > the selected one seen so far
    unsigned int* nbp;      // [n] prefix (at the width of the last selection) of the bucket to track for `next`; 0xffffffff: none
    unsigned int* dirty;    // [n] 1 if a candidate of pass 1 has non-zero low 14 bits
    unsigned int* done;     // [n] 1 once the lower middle key is known
    unsigned int* pending;  // [1] diagonals that need passes 2 and 3
};

__host__ __forceinline__ void diag_sel_bind(DiagSel& S, unsigned int* hist, unsigned int* small, size_t n, unsigned int* pending) {
    S.hist = hist;
    S.prefix = small; S.krem = small + n; S.cnt = small + 2 * n; S.need = small + 3 * n; S.next = small + 4 * n;
    S.nbp = small + 5 * n; S.dirty = small + 6 * n; S.done = small + 7 * n;
    S.pending = pending;
}

template <int PASS>
__device__ __forceinline__ void diag_hist_body(const float* __restrict__ A, long long ld, int nloc, int n, int row0, const DiagSel& S, int bx, int by) {
    const int band = by, r0 = band * DIAG_HB, r1 = min(r0 + DIAG_HB, nloc);
    const int imax = row0 + r1 - 1;
    if ((int)(bx * MV_THREADS) > imax) return;
    const int d = bx * MV_THREADS + threadIdx.x;
    if (d > imax || d >= n) return;
    if (PASS >= 2 && S.done[d]) return;
    constexpr int sh = PASS == 0 ? 23 : (PASS == 1 ? 14 : (PASS == 2 ? 7 : 0)), bits = PASS < 2 ? 9 : 7, sh_hi = sh + bits;
    constexpr unsigned int mask = (1u << bits) - 1u;
    // Everything per entry happens on the RAW float bits u: the order-preserving key is u ^ 0x80000000 for a positive and ~u for a
    // negative value, so "the high bits of the key equal the prefix" is "the high bits of u equal the prefix mapped back" (the
    // prefix fixes the sign), and the digit of the key is the digit of u, complemented for a negative prefix -- applied when a
    // slot goes to the histogram (`ordered`), not per entry.
    const unsigned int pre = PASS ? S.prefix[d] : 0u;
    const unsigned int nbpre = PASS ? S.nbp[d] : 0xffffffffu;
    constexpr unsigned int top = PASS ? (1u << ((32 - sh_hi - 1) & 31)) : 0u;           // the sign position inside a prefix
    const bool neg = PASS && !(pre & top), nbneg = PASS && !(nbpre & top);
    const unsigned int pre_raw = PASS ? (neg ? (~pre & (2u * top - 1u)) : (pre ^ top)) : 0u;
    const unsigned int nb_raw = nbpre == 0xffffffffu ? 0xffffffffu : (nbneg ? (~nbpre & (2u * top - 1u)) : (nbpre ^ top));
    const unsigned int nb_flip = nbneg ? 0xffffffffu : 0x80000000u;
    unsigned int* h = S.hist + (long long)d * DIAG_NBIN;
    const int rs = max(r0, d - row0);
    // (bin, count) slots in registers.  A batch of entries is first counted against the slots WITHOUT branches (nearly every
    // warp-row of a sparse map has both zero and non-zero lanes: a branch per entry runs both sides for everyone, and any slot
    // update inside the unrolled batch makes the compiler shuffle the slots through moves -- the first two versions of this loop
    // were issue-bound at 73 and 49 instructions per entry, profiles/r02_ncu_mcfs_batch.txt); entries that found no slot leave
    // a bit in `missed` and are handled after the batch: the slot whose turn it is goes to the histogram and is taken over.
    // The slots start on the bins short counts fall into -- pass 0: the exponents of 1, 2..3, 4..7; later passes: the digits of
    // 2^k and 3 * 2^(k-1) -- so a thread of a sparse count map never takes the slow path (a thread that starts cold misses two
    // or three times per band, and then nearly every batch of the warp has a lane in the slow path).  Any other data only pays
    // the misses it would pay anyway.
    constexpr int NS = PASS == 0 ? 3 : 2;
    unsigned int t[NS], c[NS];
#pragma unroll
    for (int k = 0; k < NS; ++k) {
        t[k] = PASS == 0 ? 127u + k : (k == 0 ? 0u : (mask + 1u) / 2u);
        c[k] = 0u;
    }
    unsigned int low = 0u, tmin = 0xffffffffu;
    int vic = 0;
    auto bin_of = [&](float v) -> unsigned int {     // the entry's (raw) bin, 0xfffffffe (a value no slot ever holds) if it does not count
        const unsigned int u = __float_as_uint(v);
        const bool nz = v != 0.f;
        if (PASS == 0) return nz ? (u >> 23) : 0xfffffffeu;
        const unsigned int hi = u >> (sh_hi & 31);
        const bool m = nz && hi == pre_raw;
        if (PASS == 1) low |= m ? u : 0u;
        tmin = min(tmin, (nz && hi == nb_raw) ? (u ^ nb_flip) : 0xffffffffu);
        return m ? ((u >> sh) & mask) : 0xfffffffeu;
    };
    auto ordered = [&](unsigned int bin) -> unsigned int {
        if (PASS == 0) return (bin & 256u) ? 511u - bin : (bin | 256u);
        return neg ? (bin ^ mask) : bin;
    };
    auto insert = [&](unsigned int bin) {            // rare: an entry whose bin is in no slot
#pragma unroll
        for (int k = 0; k < NS; ++k)
            if (bin == t[k]) { ++c[k]; return; }
#pragma unroll
        for (int k = 0; k < NS; ++k)
            if (vic == k) {
                if (c[k]) atomicAdd(h + ordered(t[k]), c[k]);
                t[k] = bin;
                c[k] = 1u;
            }
        vic = vic + 1 == NS ? 0 : vic + 1;
    };
    // the address walks down the diagonal by one row and one column per entry; the loads of the NEXT batch are in flight while
    // this one is counted (without it every warp alternated between waiting for 16 loads and ~600 instructions of counting)
    const long long stride = ld + 1;
    const float* __restrict__ p = A + ((long long)rs * stride + ((long long)row0 - d));
    int r = rs;
    float v[DIAG_BATCH];
    if (r + DIAG_BATCH <= r1) {
#pragma unroll
        for (int k = 0; k < DIAG_BATCH; ++k) { v[k] = __ldg(p); p += stride; }
    }
    for (; r + DIAG_BATCH <= r1; r += DIAG_BATCH) {
        float nv[DIAG_BATCH];
        const bool more = r + 2 * DIAG_BATCH <= r1;
        if (more) {
#pragma unroll
            for (int k = 0; k < DIAG_BATCH; ++k) { nv[k] = __ldg(p); p += stride; }
        }
        unsigned int b[DIAG_BATCH], missed = 0u;
#pragma unroll
        for (int k = 0; k < DIAG_BATCH; ++k) {
            b[k] = bin_of(v[k]);
            bool hit = false;
#pragma unroll
            for (int q = 0; q < NS; ++q) {
                const bool hq = b[k] == t[q];
                c[q] += hq;
                hit = hit || hq;
            }
            missed |= (b[k] != 0xfffffffeu && !hit) ? (1u << k) : 0u;
        }
        if (missed) {
#pragma unroll
            for (int k = 0; k < DIAG_BATCH; ++k)
                if (missed & (1u << k)) insert(b[k]);
        }
        if (more) {
#pragma unroll
            for (int k = 0; k < DIAG_BATCH; ++k) v[k] = nv[k];
        }
    }
    for (; r < r1; ++r) {
        const unsigned int bb = bin_of(__ldg(p));
        p += stride;
        if (bb != 0xfffffffeu) insert(bb);
    }
#pragma unroll
    for (int k = 0; k < NS; ++k)
        if (c[k]) atomicAdd(h + ordered(t[k]), c[k]);
    // early exit needs "all candidates have zeros in the low 14 bits of the KEY": for a positive prefix that is the low bits of
    // u; a negative prefix (its keys are ~u) simply takes all four passes
    if (PASS == 1 && (neg ? low != 0u : (low & 0x3fffu) != 0u)) S.dirty[d] = 1u;          // every writer stores the same value
    if (PASS >= 1 && tmin != 0xffffffffu) atomicMin(S.next + d, tmin);
}

template <int PASS>
__global__ void __launch_bounds__(MV_THREADS, 3)
k_diag_hist(const float* __restrict__ A, long long ld, int nloc, int n, int row0, DiagSel S) {
    if (PASS >= 2 && *S.pending == 0u) return;
    diag_hist_body<PASS>(A, ld, nloc, n, row0, S, blockIdx.x, blockIdx.y);
}

// one warp per diagonal: find the bin holding rank krem and the next occupied bin, descend, clear the histogram row.
// The row is read as uint4 groups of four bins, group j * 32 + lane by lane `lane` in round j: every request of the warp is one
// contiguous 512 bytes (lane-major bins made every load touch 32 sectors; the kernel sat on the L1 queue).  Groups that are
// already zero are not written back.
__device__ __forceinline__ void diag_select_body(int n, int pass, const DiagSel& S, int bx) {
    const int d = (bx * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (d >= n) return;
    if (pass >= 2 && S.done[d]) return;            // its histogram row was not touched in this pass
    uint4* h4 = reinterpret_cast<uint4*>(S.hist + (long long)d * DIAG_NBIN);
    constexpr int NJ = DIAG_NBIN / 128;            // rounds in the wide passes
    const int nj = (1 << diag_digit_bits(pass)) / 128;
    uint4 q[NJ];
    unsigned int gs[NJ], excl[NJ], nzm[NJ];
    unsigned int carry = 0;
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
        q[j] = make_uint4(0u, 0u, 0u, 0u);
        if (j < nj) q[j] = h4[j * 32 + lane];
    }
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
        gs[j] = q[j].x + q[j].y + q[j].z + q[j].w;
        if (j < nj && gs[j]) h4[j * 32 + lane] = make_uint4(0u, 0u, 0u, 0u);
        unsigned int incl = gs[j];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned int t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        excl[j] = carry + incl - gs[j];
        carry += __shfl_sync(0xffffffffu, incl, 31);
        nzm[j] = __ballot_sync(0xffffffffu, gs[j] != 0u);
    }
    const unsigned int total = carry;
    unsigned int kr;
    if (pass == 0) {
        if (lane == 0) { S.cnt[d] = total; S.next[d] = 0xffffffffu; }
        kr = total ? (total - 1) / 2 : 0;      // lower middle rank (np.median: mean of the two middles when even)
    } else {
        kr = S.krem[d];
    }
    if (total == 0) {
        if (lane == 0) { S.prefix[d] = 0u; S.krem[d] = 0u; S.need[d] = 0u; S.nbp[d] = 0xffffffffu; S.done[d] = 1u; }
        return;
    }
    // the group that holds rank kr (exactly one (round, lane)), known to every lane
    int jo = 0, lo = 0;
    bool mine = false;
    int first[NJ];                                 // first occupied bin of the lane's group in round j
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
        const bool mj = gs[j] != 0u && kr >= excl[j] && kr < excl[j] + gs[j];
        const unsigned int bm = __ballot_sync(0xffffffffu, mj);
        if (bm) { jo = j; lo = __ffs(bm) - 1; }
        mine = mine || mj;
        const int g4 = 4 * (j * 32 + lane);
        first[j] = q[j].x ? g4 : (q[j].y ? g4 + 1 : (q[j].z ? g4 + 2 : (q[j].w ? g4 + 3 : -1)));
    }
    // the next occupied group after the owner's: later lanes of the same round, then the following rounds
    int jt = -1, lt = 0;
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
        if (jt < 0 && j >= jo) {
            const unsigned int msk = (j == jo) ? (lo < 31 ? (nzm[j] & (0xffffffffu << (lo + 1))) : 0u) : nzm[j];
            if (msk) { jt = j; lt = __ffs(msk) - 1; }
        }
    }
    int nb_far = -1;
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
        const int f = __shfl_sync(0xffffffffu, first[j], lt);
        if (j == jt) nb_far = f;
    }
    if (mine) {
        unsigned int cq[4] = {0u, 0u, 0u, 0u};
        unsigned int run = 0;
#pragma unroll
        for (int j = 0; j < NJ; ++j)
            if (j == jo) { cq[0] = q[j].x; cq[1] = q[j].y; cq[2] = q[j].z; cq[3] = q[j].w; run = excl[j]; }
        const int g4 = 4 * (jo * 32 + lane);
        int bin = 0, nb = -1;
        unsigned int hb = 0;
        bool found = false;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            if (!found) {
                if (kr < run + cq[k]) { bin = g4 + k; hb = cq[k]; found = true; }
                else run += cq[k];
            } else if (nb < 0 && cq[k]) {
                nb = g4 + k;
            }
        }
        if (nb < 0) nb = nb_far;
        const int w = diag_digit_bits(pass), sh = diag_digit_shift(pass);
        const unsigned int pre = pass ? S.prefix[d] : 0u;
        const unsigned int sel = (pre << w) | (unsigned int)bin;
        const unsigned int kin = kr - run;      // rank inside the chosen bin
        const unsigned int nbsel = nb >= 0 ? ((pre << w) | (unsigned int)nb) : 0xffffffffu;
        const bool fin = pass == 3 || (pass == 1 && S.dirty[d] == 0u);
        if (fin) {
            S.prefix[d] = sel << sh;            // every candidate has zeros below the digits fixed so far
            const unsigned int cn = S.cnt[d];
            S.need[d] = ((cn & 1u) == 0u && kin + 1 >= hb) ? 1u : 0u;
            if (nb >= 0) S.next[d] = min(S.next[d], nbsel << sh);
            S.done[d] = 1u;
        } else {
            S.prefix[d] = sel;
            S.krem[d] = kin;
            S.nbp[d] = nbsel;
            if (pass == 0) S.done[d] = 0u;
            if (pass == 1) atomicAdd(S.pending, 1u);
        }
    }
}

__global__ void __launch_bounds__(MV_THREADS)
k_diag_select(int n, int pass, DiagSel S) {
    if (pass >= 2 && *S.pending == 0u) return;
    diag_select_body(n, pass, S, blockIdx.x);
}

__global__ void k_diag_median_final(int n, DiagSel S, double* __restrict__ avg) {
    const int d = blockIdx.x * blockDim.x + threadIdx.x;
    if (d >= n) return;
    if (S.cnt[d] == 0u) { avg[d] = 0.0; return; }                        // lib/cmstats.py:602-603
    const double lo = (double)ord2f(S.prefix[d]);
    const double hi = S.need[d] ? (double)ord2f(S.next[d]) : lo;
    avg[d] = ((S.cnt[d] & 1u) == 0u) ? (lo + hi) / 2.0 : lo;             // np.median semantics
}

// ---------------------------------------------------------------------------------------------------
// Batched distance-frequency normalization: MANY resident maps (BASELINE configs[3]: the 24 chromosomes of a genome at 25 kb)
// walked by ONE grid per stage instead of ~14 launches and two host synchronisations per map.  Every stage is the single-map
// kernel body above applied to (map, local block): a flattened block list with a per-map prefix table, so small maps ride in
// the same launch as large ones and the results are bit-identical to the per-map calls.
// ---------------------------------------------------------------------------------------------------
struct BMap {
    const float* A;
    float* Out;
    long long ld;
    int n, gx, gy, pad;
    double* avg;                 // [>= n] the map's per-diagonal statistic
    double2* avr;                // [>= n] (statistic, its reciprocal) for the apply pass
    DiagSel S;                   // median: histogram / selection state of this map
    double *psum, *sum, *cnt;    // mean: per-band partials and their reduction
    int* pcnt;
    MinMax* mm;
    int blk0;                    // first block of this map in the flattened (diagonal chunk x row band) list
    int sel0;                    // ... in the flattened one-warp-per-diagonal list
    int red0;                    // ... in the flattened 64-diagonals-per-block list
    int fin0;                    // ... in the flattened 256-diagonals-per-block list
    long long u0;                // first (band, chunk) unit of the apply pass
};

template <int FIELD>
__device__ __forceinline__ int bmap_find(const BMap* __restrict__ maps, int count, long long b) {
    // the map whose range holds b (count <= a few dozen: linear scan by every thread, uniform per CTA)
    int m = 0;
    while (m + 1 < count) {
        const BMap& nx = maps[m + 1];
        const long long start = FIELD == 0 ? nx.blk0 : (FIELD == 1 ? nx.sel0 : (FIELD == 2 ? nx.red0 : (FIELD == 3 ? nx.fin0 : nx.u0)));
        if (b < start) break;
        ++m;
    }
    return m;
}

template <int PASS>
__device__ __forceinline__ void diag_hist_block_b(const BMap* __restrict__ maps, int count, int b) {
    const int m = bmap_find<0>(maps, count, b);
    const BMap& M = maps[m];                   // fields are read where they are used: the whole table entry in registers cost occupancy
    const int local = b - M.blk0, gx = M.gx;
    diag_hist_body<PASS>(M.A, M.ld, M.n, M.n, 0, M.S, local % gx, local / gx);
}

template <int PASS>
__global__ void __launch_bounds__(MV_THREADS, 3)
k_diag_hist_b(const BMap* __restrict__ maps, int count, int nblocks) {
    if (PASS < 2) {
        diag_hist_block_b<PASS>(maps, count, blockIdx.x);          // one CTA per block of the list
    } else {
        // passes 2 and 3 run on a capped grid that strides over the block list: when no diagonal is pending (count maps) they
        // cost one wave of empty CTAs
        if (*maps[0].S.pending == 0u) return;
        for (int b = blockIdx.x; b < nblocks; b += gridDim.x) diag_hist_block_b<PASS>(maps, count, b);
    }
}

__global__ void __launch_bounds__(MV_THREADS)
k_diag_select_b(const BMap* __restrict__ maps, int count, int pass, int nblocks) {
    if (pass >= 2 && *maps[0].S.pending == 0u) return;
    for (int b = blockIdx.x; b < nblocks; b += gridDim.x) {
        const int m = bmap_find<1>(maps, count, b);
        const BMap& M = maps[m];
        diag_select_body(M.n, pass, M.S, b - M.sel0);
    }
}

__global__ void __launch_bounds__(256)
k_diag_median_final_b(const BMap* __restrict__ maps, int count) {
    const int m = bmap_find<3>(maps, count, blockIdx.x);
    const BMap M = maps[m];
    const int d = ((int)blockIdx.x - M.fin0) * 256 + threadIdx.x;
    if (d >= M.n) return;
    if (M.S.cnt[d] == 0u) { M.avg[d] = 0.0; return; }
    const double lo = (double)ord2f(M.S.prefix[d]);
    const double hi = M.S.need[d] ? (double)ord2f(M.S.next[d]) : lo;
    M.avg[d] = ((M.S.cnt[d] & 1u) == 0u) ? (lo + hi) / 2.0 : lo;
}

__global__ void __launch_bounds__(MV_THREADS)
k_diag_mean_partial_b(const BMap* __restrict__ maps, int count) {
    const int m = bmap_find<0>(maps, count, blockIdx.x);
    const BMap M = maps[m];
    const int local = (int)blockIdx.x - M.blk0;
    diag_mean_partial_body(M.A, M.ld, M.n, M.n, 0, M.psum, M.pcnt, local % M.gx, local / M.gx);
}

__global__ void __launch_bounds__(64 * DIAG_RED_SL)
k_diag_mean_reduce_b(const BMap* __restrict__ maps, int count) {
    const int m = bmap_find<2>(maps, count, blockIdx.x);
    const BMap M = maps[m];
    diag_mean_reduce_body(M.gy, M.n, 0, M.psum, M.pcnt, M.sum, M.cnt, (int)blockIdx.x - M.red0);
}

__global__ void __launch_bounds__(256)
k_diag_mean_final_b(const BMap* __restrict__ maps, int count) {
    const int m = bmap_find<3>(maps, count, blockIdx.x);
    const BMap M = maps[m];
    const int d = ((int)blockIdx.x - M.fin0) * 256 + threadIdx.x;
    if (d < M.n) M.avg[d] = M.cnt[d] > 0.0 ? M.sum[d] / M.cnt[d] : 0.0;
}

// pooled mean over several maps (lib/cmstats.py:572-578 with a list of maps): sum of the maps' per-diagonal sums / counts
__global__ void __launch_bounds__(256)
k_diag_pool_mean_final(const BMap* __restrict__ maps, int count, int nmax, double* __restrict__ avg) {
    const int d = blockIdx.x * blockDim.x + threadIdx.x;
    if (d >= nmax) return;
    double s = 0.0, c = 0.0;
    for (int m = 0; m < count; ++m) {
        if (d < maps[m].n) { s += maps[m].sum[d]; c += maps[m].cnt[d]; }
    }
    avg[d] = c > 0.0 ? s / c : 0.0;
}

__global__ void k_mm_init_b(const BMap* __restrict__ maps, int count) {
    const int m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m < count) { maps[m].mm->min_nz = 0xffffffffu; maps[m].mm->max_all = 0u; }
}
__global__ void k_mm_gather_b(const BMap* __restrict__ maps, int count, MinMax* out) {
    const int m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m < count) out[m] = *maps[m].mm;
}

__global__ void __launch_bounds__(256)
k_diag_recip_b(const BMap* __restrict__ maps, int count) {
    const int m = bmap_find<3>(maps, count, blockIdx.x);
    const BMap& M = maps[m];
    const int d = ((int)blockIdx.x - M.fin0) * 256 + threadIdx.x;
    if (d < M.n) {
        const double a = M.avg[d];
        M.avr[d] = make_double2(a, a != 0.0 ? 1.0 / a : 0.0);
    }
}

// the apply pass of all maps: the (band, chunk) units of every map in one list, split evenly over a persistent grid
template <int R, int OCC>
__global__ void __launch_bounds__(MV_THREADS, OCC)
k_diag_apply_b(const BMap* __restrict__ maps, int count, long long U, int subtract) {
    const int tid = threadIdx.x;
    const long long G = gridDim.x, c = blockIdx.x;
    const long long ua = unit_begin(c, U, G), ub = unit_begin(c + 1, U, G);
    if (ua >= ub) return;
    int m = bmap_find<4>(maps, count, ua);
    long long u = ua;
    while (u < ub) {
        const BMap M = maps[m];
        const long long uend = (m + 1 < count) ? maps[m + 1].u0 : U;
        const long long ue = uend < ub ? uend : ub;
        const int n = M.n, n4 = (int)(M.ld / 4);
        const int nch = (n4 + MV_THREADS - 1) / MV_THREADS;
        const double2* __restrict__ avr = M.avr;
        float mn = GCX_PINF, mx = GCX_NINF;
        for (; u < ue; ++u) {
            const long long lu = u - M.u0;
            const int b = (int)(lu / nch), k = (int)(lu - (long long)b * nch);
            const int c4 = k * MV_THREADS + tid;
            if (c4 >= n4) continue;
            float4 a[R];
#pragma unroll
            for (int rr = 0; rr < R; ++rr) {
                const int r = b * R + rr;
                if (r < n) a[rr] = ld_stream_f4(reinterpret_cast<const float4*>(M.A + (long long)r * M.ld) + c4);
            }
            const int j0 = 4 * c4;
#pragma unroll
            for (int rr = 0; rr < R; ++rr) {
                const int i = b * R + rr;
                if (i >= n) break;
                const float in[4] = {a[rr].x, a[rr].y, a[rr].z, a[rr].w};
                float o[4];
                if (subtract) diag_apply_quad<1, true, true>(in, o, i, j0, n, avr, mn, mx);      // lib/normalizeCore.pyx:138-144
                else diag_apply_quad<1, false, true>(in, o, i, j0, n, avr, mn, mx);              // :105-117
                st_stream_f4(reinterpret_cast<float4*>(M.Out + (long long)i * M.ld) + c4, make_float4(o[0], o[1], o[2], o[3]));
            }
        }
        mm_commit(mn, mx, M.mm);
        ++m;
    }
}

// The same pass with lane-strided columns: lane l of warp w takes columns 128 w + l, + 32, + 64, + 96 of the unit's 1024.  Every
// load, store and table lookup of a warp then covers CONSECUTIVE addresses (128 bytes of the map, 512 bytes of the (avg, 1/avg)
// table); with a float4 per thread the four lookups of a warp each touched all 16 lines of a 2 KB window and the L1 ran at 87 %
// (profiles/r02_ncu_mcfs_batch.txt).
__device__ __forceinline__ float ld_stream_f1(const float* p) {
    float r;
    asm("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(r) : "l"(p));
    return r;
}
__device__ __forceinline__ void st_stream_f1(float* p, float v) {
    asm volatile("st.global.L1::no_allocate.f32 [%0], %1;" ::"l"(p), "f"(v) : "memory");
}

// one map (possibly a row block [row0, row0 + nloc) of a sharded one)
template <int R, int OCC, bool SUB>
__global__ void __launch_bounds__(MV_THREADS, OCC)
k_diag_apply_lane(const float* __restrict__ A, float* __restrict__ Out, long long ld, int nloc, int n, int row0,
                  const double2* __restrict__ avr, MinMax* mm) {
    const int tid = threadIdx.x, cl = (tid >> 5) * 128 + (tid & 31);
    const int nch = ((int)(ld / 4) + MV_THREADS - 1) / MV_THREADS;
    const int nb = (nloc + R - 1) / R;
    const long long U = (long long)nb * nch, G = gridDim.x, c = blockIdx.x;
    const long long u0 = unit_begin(c, U, G), u1 = unit_begin(c + 1, U, G);
    float mn = GCX_PINF, mx = GCX_NINF;
    for (long long u = u0; u < u1; ++u) {
        const int b = (int)(u / nch), k = (int)(u - (long long)b * nch);
        const int j0 = k * 1024 + cl, r0 = b * R;
        const float* __restrict__ ap = A + (long long)r0 * ld + j0;
        float* __restrict__ op = Out + (long long)r0 * ld + j0;
        float a[R][4];
        if (k * 1024 + 1024 <= n && r0 + R <= nloc) {
#pragma unroll
            for (int rr = 0; rr < R; ++rr)
#pragma unroll
                for (int e = 0; e < 4; ++e) a[rr][e] = ld_stream_f1(ap + rr * ld + 32 * e);
#pragma unroll
            for (int rr = 0; rr < R; ++rr) {
                float o[4];
                diag_apply_quad<32, SUB, false>(a[rr], o, row0 + r0 + rr, j0, n, avr, mn, mx);
#pragma unroll
                for (int e = 0; e < 4; ++e) st_stream_f1(op + rr * ld + 32 * e, o[e]);
            }
        } else {
#pragma unroll
            for (int rr = 0; rr < R; ++rr)
#pragma unroll
                for (int e = 0; e < 4; ++e) a[rr][e] = (r0 + rr < nloc && j0 + 32 * e < (int)ld) ? ld_stream_f1(ap + rr * ld + 32 * e) : 0.f;
#pragma unroll
            for (int rr = 0; rr < R; ++rr) {
                if (r0 + rr >= nloc) break;
                float o[4];
                diag_apply_quad<32, SUB, true>(a[rr], o, row0 + r0 + rr, j0, n, avr, mn, mx);
#pragma unroll
                for (int e = 0; e < 4; ++e)
                    if (j0 + 32 * e < (int)ld) st_stream_f1(op + rr * ld + 32 * e, o[e]);
            }
        }
    }
    mm_commit(mn, mx, mm);
}

template <int R, int OCC, bool SUB>
__global__ void __launch_bounds__(MV_THREADS, OCC)
k_diag_apply_lane_b(const BMap* __restrict__ maps, int count, long long U) {
    const int tid = threadIdx.x, cl = (tid >> 5) * 128 + (tid & 31);
    const long long G = gridDim.x, c = blockIdx.x;
    const long long ua = unit_begin(c, U, G), ub = unit_begin(c + 1, U, G);
    if (ua >= ub) return;
    int m = bmap_find<4>(maps, count, ua);
    long long u = ua;
    while (u < ub) {
        const BMap& M = maps[m];
        const long long uend = (m + 1 < count) ? maps[m + 1].u0 : U, mu0 = M.u0;
        const long long ue = uend < ub ? uend : ub;
        const int n = M.n;
        const long long ld = M.ld;
        const int nch = ((int)(ld / 4) + MV_THREADS - 1) / MV_THREADS;
        const double2* __restrict__ avr = M.avr;
        const float* __restrict__ A = M.A;
        float* __restrict__ Out = M.Out;
        float mn = GCX_PINF, mx = GCX_NINF;
        for (; u < ue; ++u) {
            const long long lu = u - mu0;
            const int b = (int)(lu / nch), k = (int)(lu - (long long)b * nch);
            const int j0 = k * 1024 + cl, i0 = b * R;
            const float* __restrict__ ap = A + (long long)i0 * ld + j0;
            float* __restrict__ op = Out + (long long)i0 * ld + j0;
            float a[R][4];
            if (k * 1024 + 1024 <= n && i0 + R <= n) {
                // interior unit: every row and column exists
#pragma unroll
                for (int rr = 0; rr < R; ++rr)
#pragma unroll
                    for (int e = 0; e < 4; ++e) a[rr][e] = ld_stream_f1(ap + rr * ld + 32 * e);
#pragma unroll
                for (int rr = 0; rr < R; ++rr) {
                    float o[4];
                    diag_apply_quad<32, SUB, false>(a[rr], o, i0 + rr, j0, n, avr, mn, mx);       // lib/normalizeCore.pyx:105-117, :138-144
#pragma unroll
                    for (int e = 0; e < 4; ++e) st_stream_f1(op + rr * ld + 32 * e, o[e]);
                }
            } else {
#pragma unroll
                for (int rr = 0; rr < R; ++rr)
#pragma unroll
                    for (int e = 0; e < 4; ++e) a[rr][e] = (i0 + rr < n && j0 + 32 * e < (int)ld) ? ld_stream_f1(ap + rr * ld + 32 * e) : 0.f;
#pragma unroll
                for (int rr = 0; rr < R; ++rr) {
                    if (i0 + rr >= n) break;
                    float o[4];
                    diag_apply_quad<32, SUB, true>(a[rr], o, i0 + rr, j0, n, avr, mn, mx);
#pragma unroll
                    for (int e = 0; e < 4; ++e)
                        if (j0 + 32 * e < (int)ld) st_stream_f1(op + rr * ld + 32 * e, o[e]);
                }
            }
        }
        mm_commit(mn, mx, M.mm);
        ++m;
    }
}

// ---------------------------------------------------------------------------------------------------
// Coarsening: downSample2DMap, lib/ccmap.py:737-842.  One thread per output element.
// Output row/column 0 stay zero; output (c, 1+g) reduces input rows 1+(c-1)*level.. and columns 1+g*level..
// The reduction order is NumPy's, so results are bit-identical: over the rows of the block sequentially, then over
// the `level` column results with NumPy's pairwise_sum (sequential below 8 terms; 8 interleaved accumulators, a
// fixed combine tree and a sequential tail up to 128 terms).  When the column count needs padding the reference
// hstacks float64 zeros, which promotes that whole case to float64 arithmetic (T = double); otherwise T = float.
// method: 0 sum, 1 mean, 2 max.
// ---------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256)
k_downsample(const float* __restrict__ A, long long ld, int n, int level, int method, int on, float* __restrict__ out, MinMax* mm) {
    float mn = GCX_PINF, mx = GCX_NINF;
    const long long total = (long long)on * on;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(idx / on), gq = (int)(idx - (long long)c * on);
        float res = 0.f;
        if (c > 0 && gq > 0) {
            const int i0 = 1 + (c - 1) * level, j0 = 1 + (gq - 1) * level;
            const int rows = min(level, n - i0);
            T r8[8];
            T acc = (T)0;
            const int nfull = level - (level % 8);
            for (int j = 0; j < level; ++j) {
                // column j of the block, reduced over its rows (missing columns are the zero padding)
                T cv = (T)0;
                const int col = j0 + j;
                if (col < n) {
                    if (method == 2) {
                        cv = (T)A[(long long)i0 * ld + col];
                        for (int r = 1; r < rows; ++r) cv = max(cv, (T)A[(long long)(i0 + r) * ld + col]);
                    } else {
                        for (int r = 0; r < rows; ++r) cv = cv + (T)A[(long long)(i0 + r) * ld + col];
                        if (method == 1) cv = cv / (T)rows;
                    }
                }
                // feed NumPy's reduction over the `level` column results
                if (method == 2) {
                    acc = (j == 0) ? cv : max(acc, cv);
                } else if (level < 8) {
                    acc = acc + cv;
                } else if (j < 8) {
                    r8[j] = cv;
                } else if (j < nfull) {
                    r8[j & 7] = r8[j & 7] + cv;
                } else {
                    if (j == nfull) acc = ((r8[0] + r8[1]) + (r8[2] + r8[3])) + ((r8[4] + r8[5]) + (r8[6] + r8[7]));
                    acc = acc + cv;
                }
            }
            if (method != 2 && level >= 8 && nfull == level) acc = ((r8[0] + r8[1]) + (r8[2] + r8[3])) + ((r8[4] + r8[5]) + (r8[6] + r8[7]));
            if (method == 1) acc = acc / (T)level;
            res = (float)acc;
        }
        out[idx] = res;
        mm_update(res, mn, mx);
    }
    mm_commit(mn, mx, mm);
}

// K0 compaction: out[ci][cj] = (double) A[kept[ci]][kept[cj]]  -- ccmapHelpers.remove_zeros (lib/ccmapHelpers.pyx:286-299),
// which builds a float64 matrix of the kept rows/columns.  The normalizers here never compact (masked bins are zeroed vector
// entries); this serves callers of the inner seam.
__global__ void __launch_bounds__(256)
k_compact_f64(const float* __restrict__ A, long long ld, const int* __restrict__ kept_rows, int nrows, const int* __restrict__ kept_cols,
              int nk, double* __restrict__ out) {
    for (int ci = blockIdx.x; ci < nrows; ci += gridDim.x) {
        const float* row = A + (long long)kept_rows[ci] * ld;
        for (int cj = threadIdx.x; cj < nk; cj += blockDim.x) out[(long long)ci * nk + cj] = (double)row[kept_cols[cj]];
    }
}

// vmin/vmax clamp + scale, in place (lib/normalizer.py:567-569, 832-834)
__global__ void __launch_bounds__(MV_THREADS)
k_clamp_scale(float* __restrict__ A, long long total4, int has_vmin, float vmin, int has_vmax, float vmax, int has_scale,
              float scale) {
    float4* p = reinterpret_cast<float4*>(A);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total4; i += (long long)gridDim.x * blockDim.x) {
        float4 v = p[i];
        float e[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            if (has_vmin && e[k] <= vmin) e[k] = 0.f;
            if (has_vmax && e[k] >= vmax) e[k] = 0.f;
            if (has_scale) e[k] = e[k] * scale;
        }
        p[i] = make_float4(e[0], e[1], e[2], e[3]);
    }
}

// ---------------------------------------------------------------------------------------------------
// synthetic map: value is a pure function of (seed, min(i,j), max(i,j)) so any row block
// can be generated on any rank.  Poisson(lambda(d)), lambda = scale * (d+1)^-alpha.
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ bool synth_empty(uint64_t seed, int i, unsigned int thresh24) {
    return (unsigned int)(mix64(seed * 0x2545F4914F6CDD1Dull + 0x5851F42D4C957F2Dull * (uint64_t)(i + 1)) >> 40) < thresh24;
}

__device__ __forceinline__ float synth_value(uint64_t seed, int i, int j, float scale, float alpha, unsigned int thresh24) {
    const int lo = i < j ? i : j, hi = i < j ? j : i;
    if (synth_empty(seed, lo, thresh24) || synth_empty(seed, hi, thresh24)) return 0.f;
    const uint64_t h = mix64(mix64(seed ^ ((uint64_t)lo << 32 | (uint64_t)hi)) + seed);
    const float u1 = ((unsigned int)(h >> 40) + 0.5f) * (1.0f / 16777216.0f);
    const float u2 = ((unsigned int)((h >> 16) & 0xffffffu) + 0.5f) * (1.0f / 16777216.0f);
    const float lam = scale * __powf((float)(hi - lo + 1), -alpha);
    float kf;
    if (lam < 24.f) {
        float p = __expf(-lam), F = p;
        int k = 0;
        while (u1 > F && k < 128) {
            ++k;
            p *= lam / (float)k;
            F += p;
        }
        kf = (float)k;
    } else {
        const float zn = sqrtf(-2.f * __logf(u1)) * __cosf(6.2831853f * u2);
        kf = floorf(lam + sqrtf(lam) * zn + 0.5f);
        if (kf < 0.f) kf = 0.f;
    }
    if (lo == hi && kf < 1.f) kf = 1.f;
    return kf;
}

__global__ void __launch_bounds__(MV_THREADS)
k_synth(float* __restrict__ A, long long ld, int nloc, int n, int row0, uint64_t seed, float scale, float alpha,
        unsigned int thresh24) {
    const int n4 = (int)(ld / 4);
    for (int r = blockIdx.x; r < nloc; r += gridDim.x) {
        const int i = row0 + r;
        float4* rp = reinterpret_cast<float4*>(A + (long long)r * ld);
        for (int c4 = threadIdx.x; c4 < n4; c4 += MV_THREADS) {
            float e[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int j = 4 * c4 + k;
                e[k] = j < n ? synth_value(seed, i, j, scale, alpha, thresh24) : 0.f;
            }
            rp[c4] = make_float4(e[0], e[1], e[2], e[3]);
        }
    }
}

// order-independent checksum: sum over entries of mix64(bits ^ (i*n+j)) (sharded parity checks)
__global__ void __launch_bounds__(MV_THREADS)
k_checksum(const float* __restrict__ A, long long ld, int nloc, int n, int row0, unsigned long long* out) {
    unsigned long long acc = 0;
    for (int r = blockIdx.x; r < nloc; r += gridDim.x) {
        const float* rp = A + (long long)r * ld;
        const uint64_t base = (uint64_t)(row0 + r) * (uint64_t)n;
        for (int j = threadIdx.x; j < n; j += MV_THREADS) {
            const unsigned int bits = __float_as_uint(rp[j]);
            if (bits & 0x7fffffffu) acc += mix64((uint64_t)bits ^ ((base + j) << 32));
        }
    }
    acc = (unsigned long long)warp_sum_ll((long long)acc);
    if ((threadIdx.x & 31) == 0) atomicAdd(out, acc);
}

// ---------------------------------------------------------------------------------------------------
// IC vector epilogue -- the O(n) part of lib/normalizeCore.pyx:42-60, one launch per matvec.
//
// With w = 1/B_k the reference's pass k+1 is: dBraw = w o (A w); dB = dBraw/mean(dBraw[!=0]) (0 -> 1);
// v = w/dB; S' = v'Av; B_{k+1} = B_k dB sqrt(S'/cc); next dBraw = (cc/S') v o (A v).  So matvec m (t = A v)
// closes pass m and opens pass m+1.  Single CTA, fixed reduction tree -> every rank that
// runs this on the same gathered vector gets bitwise identical state.
// ---------------------------------------------------------------------------------------------------
struct IcState {
    double cc, S, l1, mean, tol;
    int pass, done, cnt, max_iter, matvecs;
    // deferred-scalar pass (k_ic_pass_symd): v = alpha * u for the vector u the next product multiplies, B = beta / u on bins with
    // data and phi on kept bins without, for the last closed pass; irregular / error make the host fall back / fail
    int irregular, error;
    double alpha, beta, phi;
};

__global__ void __launch_bounds__(EP_THREADS)
k_ic_init(IcState* st, int nv, const uint8_t* __restrict__ keep, double* v, double* B, double tol, int max_iter) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nv; i += gridDim.x * blockDim.x) {
        v[i] = keep[i] ? 1.0 : 0.0;
        B[i] = 1.0;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        st->cc = 0; st->S = 0; st->l1 = 0; st->mean = 0; st->tol = tol;
        st->pass = 0; st->done = 0; st->cnt = 0; st->max_iter = max_iter; st->matvecs = 0;
        st->irregular = 0; st->error = 0; st->alpha = 1.0; st->beta = 1.0; st->phi = 1.0;
    }
}

// State per bin: B_k and v = 1/(B_k dB_{k+1}) (the matvec input).  Two phases per launch:
//   A: S' = sum v_i t_i and the number of non-zero terms            (one cluster reduction)
//   B: B_{k+1} = f / v  (= B_k dB f, :51-52), L1 change (:56), next v (one cluster reduction for the L1)
// mean(dBraw[!=0]) (:44) of the next pass is (cc/S') S' / cnt -- the sum of the raw column sums is S' itself.
__global__ void __cluster_dims__(EP_CL, 1, 1) __launch_bounds__(EP_THREADS)
k_ic_epilogue(IcState* st, int n, const uint8_t* __restrict__ keep, const double* __restrict__ t, double* v, double* B) {
    __shared__ double sm[32];
    __shared__ double slots[8];
    if (st->done) return;
    Team T(slots, sm);
    const int m = st->matvecs;
    const double cc_prev = st->cc, tol = st->tol;
    const int max_iter = st->max_iter;
    double a[2] = {0.0, 0.0};
    for (int i = T.gtid; i < n; i += T.nthreads) {
        if (keep[i]) {
            const double p = v[i] * t[i];
            a[0] += p;
            a[1] += (p != 0.0) ? 1.0 : 0.0;
        }
    }
    T.reduce<RED_SUM, 2>(a);
    const double S = a[0], C = a[1];
    const double cc = (m == 0) ? S : cc_prev;           // contact_count, :36
    const double f = (m == 0) ? 1.0 : sqrt(S / cc);     // :52
    const double g = (m == 0) ? 1.0 : cc / S;           // :53
    const double mean = g * S / C;                      // :44
    double l1 = 0.0;
    for (int i = T.gtid; i < n; i += T.nthreads) {
        if (keep[i]) {
            const double vi = v[i];
            const double d = g * (vi * t[i]);           // next pass's raw column sum, :42
            double bn = B[i];
            if (m > 0) {
                bn = f / vi;                            // (B dB) sqrt(S'/cc), :51-52
                l1 += fabs(B[i] - bn);                  // :56
                B[i] = bn;
            }
            v[i] = (d != 0.0) ? mean / (bn * d) : 1.0 / bn;   // 1/(B dB), dB = d/mean or 1 (:44-45)
        }
    }
    const double L1 = T.reduce1<RED_SUM>(l1);
    const bool done = (m >= 2) && (L1 < tol || (m - 1) > max_iter);   // :55-57 (count = pass-1)
    if (T.gtid == 0) {
        st->cc = cc; st->S = S; st->l1 = L1; st->mean = mean; st->cnt = (int)C;
        st->pass = m; st->matvecs = m + 1; st->done = done ? 1 : 0;
    }
    T.finish();
}

// ---------------------------------------------------------------------------------------------------
// Fused IC pass: ONE cooperative launch per iteration = [vector step that closes pass m-1] + [matvec m].
//   P1  p_i = v_i t_i ; per-CTA partial (S', count)                      -> grid barrier
//   P2  every CTA folds the partials in CTA order (bitwise identical scalars everywhere), then for its
//       bins: B_{k+1} = f / v, L1 change, next v ; per-CTA partial L1     -> grid barrier (v complete)
//   P3  fold the L1 partials; converged -> record state, whole grid exits
//   P4  the matvec main loop of k_matvec with z = v ; last CTA folds split bands and bumps the counter
// Multi-GPU: the host enqueues one NCCL allgather of t after each launch; P1..P3 then run redundantly on
// every rank on the same gathered vector with the same grid, so all ranks stay bitwise identical.
// `part` holds 3 x gridDim.x doubles.
// ---------------------------------------------------------------------------------------------------
template <int R, int OCC>
__global__ void __launch_bounds__(MV_THREADS, OCC)
k_ic_pass(const float* __restrict__ A, long long ld, int nloc, int n4, int n, int row0, const uint8_t* __restrict__ keep,
          double* t, double* v, double* B, double* ws, unsigned int* ticket, IcState* st, double* part) {
    namespace cg = cooperative_groups;
    if (st->done) return;
    cg::grid_group grid = cg::this_grid();
    __shared__ double sm[32];
    __shared__ double red[MV_THREADS / 32][R];
    __shared__ int s_last;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int G = gridDim.x, bid = blockIdx.x;
    const int m = st->matvecs;                  // matvecs completed so far; t holds the result of matvec m-1
    const int gth = bid * MV_THREADS + tid, nth = G * MV_THREADS;

    if (m > 0) {
        const int mm = m - 1;
        const double cc_prev = st->cc, tol = st->tol;
        const int max_iter = st->max_iter;
        // ---- P1 ----
        double s = 0.0, c = 0.0;
        for (int i = gth; i < n; i += nth) {
            if (keep[i]) {
                const double p = v[i] * t[i];
                s += p;
                c += (p != 0.0) ? 1.0 : 0.0;
            }
        }
        s = block_reduce<RED_SUM>(s, sm);
        c = block_reduce<RED_SUM>(c, sm);
        if (tid == 0) {
            part[bid] = s;
            part[G + bid] = c;
        }
        grid.sync();
        // ---- P2 ----
        double a0 = 0.0, a1 = 0.0;
        for (int b = tid; b < G; b += MV_THREADS) {
            a0 += __ldcg(part + b);
            a1 += __ldcg(part + G + b);
        }
        const double S = block_reduce<RED_SUM>(a0, sm), C = block_reduce<RED_SUM>(a1, sm);
        const double cc = (mm == 0) ? S : cc_prev;          // contact_count, lib/normalizeCore.pyx:36
        const double f = (mm == 0) ? 1.0 : sqrt(S / cc);    // :52
        const double g = (mm == 0) ? 1.0 : cc / S;          // :53
        const double mean = g * S / C;                      // :44
        double l1 = 0.0;
        for (int i = gth; i < n; i += nth) {
            if (keep[i]) {
                const double vi = v[i];
                const double d = g * (vi * t[i]);           // next pass's raw column sum, :42
                double bn = B[i];
                if (mm > 0) {
                    bn = f / vi;                            // (B dB) sqrt(S'/cc), :51-52
                    l1 += fabs(B[i] - bn);                  // :56
                    B[i] = bn;
                }
                v[i] = (d != 0.0) ? mean / (bn * d) : 1.0 / bn;     // 1/(B dB), :44-45
            }
        }
        l1 = block_reduce<RED_SUM>(l1, sm);
        if (tid == 0) part[2 * G + bid] = l1;
        grid.sync();
        // ---- P3 ----
        double a2 = 0.0;
        for (int b = tid; b < G; b += MV_THREADS) a2 += __ldcg(part + 2 * G + b);
        const double L1 = block_reduce<RED_SUM>(a2, sm);
        const bool done = (mm >= 2) && (L1 < tol || (mm - 1) > max_iter);       // :55-57
        if (bid == 0 && tid == 0) {
            st->cc = cc; st->S = S; st->l1 = L1; st->mean = mean; st->cnt = (int)C;
            st->pass = mm; st->done = done ? 1 : 0;
        }
        if (done) return;
    }

    // ---- P4: matvec (same decomposition and arithmetic as k_matvec; z = v is read with coherent loads
    //      because this kernel wrote it before the grid barrier) ----
    const int nch = (n4 + MV_THREADS - 1) / MV_THREADS;
    const int nb = (nloc + R - 1) / R;
    const long long U = (long long)nb * nch, GG = G, cta = bid;
    long long u = unit_begin(cta, U, GG);
    const long long u1 = unit_begin(cta + 1, U, GG);
    const int bfirst = (int)(u / nch);
    const double2* z2 = reinterpret_cast<const double2*>(v);
    double* q = t + row0;
    while (u < u1) {
        const int b = (int)(u / nch);
        const int k0 = (int)(u - (long long)b * nch);
        const long long rem = u1 - u;
        const int k1 = (int)(((long long)(nch - k0) < rem) ? nch : (k0 + rem));
        const float4* rowp[R];
#pragma unroll
        for (int rr = 0; rr < R; ++rr) {
            int r = b * R + rr;
            r = r < nloc ? r : nloc - 1;
            rowp[rr] = reinterpret_cast<const float4*>(A + (long long)r * ld);
        }
        double acc[R];
#pragma unroll
        for (int rr = 0; rr < R; ++rr) acc[rr] = 0.0;
        for (int k = k0; k < k1; ++k) {
            const int c4 = k * MV_THREADS + tid;
            if (c4 < n4) {
                float4 a[R];
#pragma unroll
                for (int rr = 0; rr < R; ++rr) a[rr] = ld_stream_f4(rowp[rr] + c4);
                const double2 za = z2[2 * c4], zb = z2[2 * c4 + 1];
#pragma unroll
                for (int rr = 0; rr < R; ++rr) {
                    acc[rr] = fma((double)a[rr].x, za.x, acc[rr]);
                    acc[rr] = fma((double)a[rr].y, za.y, acc[rr]);
                    acc[rr] = fma((double)a[rr].z, zb.x, acc[rr]);
                    acc[rr] = fma((double)a[rr].w, zb.y, acc[rr]);
                }
            }
        }
#pragma unroll
        for (int rr = 0; rr < R; ++rr) acc[rr] = warp_sum(acc[rr]);
        if (lane == 0) {
#pragma unroll
            for (int rr = 0; rr < R; ++rr) red[warp][rr] = acc[rr];
        }
        __syncthreads();
        if (tid < R) {
            double s = 0.0;
#pragma unroll
            for (int w = 0; w < MV_THREADS / 32; ++w) s += red[w][tid];
            const int r = b * R + tid;
            if (k0 == 0 && k1 == nch) {
                if (r < nloc) q[r] = s;
            } else {
                ws[((long long)cta * 2 + (b == bfirst ? 0 : 1)) * R + tid] = s;
            }
        }
        __syncthreads();
        u += (k1 - k0);
    }
    __threadfence();
    __syncthreads();
    if (tid == 0) s_last = (atomicAdd(ticket, 1u) == (unsigned int)(G - 1));
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    for (long long cb = 1 + tid; cb < GG; cb += MV_THREADS) {
        const long long ub = unit_begin(cb, U, GG);
        const int b = (int)(ub / nch);
        const long long band0 = (long long)b * nch, band1 = band0 + nch;
        if (ub == band0 || ub >= U) continue;
        const long long ubp = unit_begin(cb - 1, U, GG);
        if (ubp > band0) continue;
        double s[R];
        {
            const int slot = (b == (int)(ubp / nch)) ? 0 : 1;
#pragma unroll
            for (int rr = 0; rr < R; ++rr) s[rr] = __ldcg(ws + ((cb - 1) * 2 + slot) * R + rr);
        }
        long long cc2 = cb, ucc = ub;
        while (cc2 < GG && ucc < band1) {
            const long long unext = unit_begin(cc2 + 1, U, GG);
            if (unext > ucc) {
#pragma unroll
                for (int rr = 0; rr < R; ++rr) s[rr] += __ldcg(ws + (cc2 * 2) * R + rr);
            }
            ++cc2;
            ucc = unext;
        }
#pragma unroll
        for (int rr = 0; rr < R; ++rr) {
            const int r = b * R + rr;
            if (r < nloc) q[r] = s[rr];
        }
    }
    if (tid == 0) {
        *ticket = 0u;
        st->matvecs = m + 1;
    }
}

// Fused IC pass with the TMA-bulk main loop (same P1..P4 structure as k_ic_pass; all 288 threads take part in
// the vector phases).
template <int R, int STAGES, int OCC>
__global__ void __launch_bounds__(TMA_THREADS, OCC)
k_ic_pass_tma(const float* __restrict__ A, long long ld, int nloc, int n4, int n, int row0, const uint8_t* __restrict__ keep, double* t,
              double* v, double* B, double* ws, unsigned int* ticket, IcState* st, double* part) {
    namespace cg = cooperative_groups;
    if (st->done) return;
    cg::grid_group grid = cg::this_grid();
    extern __shared__ __align__(128) unsigned char tma_smem[];
    __shared__ __align__(8) uint64_t full_bar[STAGES];
    __shared__ __align__(8) uint64_t empty_bar[STAGES];
    __shared__ double red[MV_THREADS / 32][R];
    __shared__ double sm[32];
    __shared__ int s_last;
    const int tid = threadIdx.x;
    const int G = gridDim.x, bid = blockIdx.x;
    const int m = st->matvecs;
    const int gth = bid * TMA_THREADS + tid, nth = G * TMA_THREADS;

    if (m > 0) {
        const int mm = m - 1;
        const double cc_prev = st->cc, tol = st->tol;
        const int max_iter = st->max_iter;
        double s = 0.0, c = 0.0;
        for (int i = gth; i < n; i += nth) {
            if (keep[i]) {
                const double p = v[i] * t[i];
                s += p;
                c += (p != 0.0) ? 1.0 : 0.0;
            }
        }
        s = block_reduce<RED_SUM>(s, sm);
        c = block_reduce<RED_SUM>(c, sm);
        if (tid == 0) {
            part[bid] = s;
            part[G + bid] = c;
        }
        grid.sync();
        double a0 = 0.0, a1 = 0.0;
        for (int b = tid; b < G; b += TMA_THREADS) {
            a0 += __ldcg(part + b);
            a1 += __ldcg(part + G + b);
        }
        const double S = block_reduce<RED_SUM>(a0, sm), C = block_reduce<RED_SUM>(a1, sm);
        const double cc = (mm == 0) ? S : cc_prev;          // contact_count, lib/normalizeCore.pyx:36
        const double f = (mm == 0) ? 1.0 : sqrt(S / cc);    // :52
        const double g = (mm == 0) ? 1.0 : cc / S;          // :53
        const double mean = g * S / C;                      // :44
        double l1 = 0.0;
        for (int i = gth; i < n; i += nth) {
            if (keep[i]) {
                const double vi = v[i];
                const double d = g * (vi * t[i]);
                double bn = B[i];
                if (mm > 0) {
                    bn = f / vi;                            // :51-52
                    l1 += fabs(B[i] - bn);                  // :56
                    B[i] = bn;
                }
                v[i] = (d != 0.0) ? mean / (bn * d) : 1.0 / bn;     // :44-45
            }
        }
        l1 = block_reduce<RED_SUM>(l1, sm);
        if (tid == 0) part[2 * G + bid] = l1;
        grid.sync();
        double a2 = 0.0;
        for (int b = tid; b < G; b += TMA_THREADS) a2 += __ldcg(part + 2 * G + b);
        const double L1 = block_reduce<RED_SUM>(a2, sm);
        const bool done = (mm >= 2) && (L1 < tol || (mm - 1) > max_iter);       // :55-57
        if (bid == 0 && tid == 0) {
            st->cc = cc; st->S = S; st->l1 = L1; st->mean = mean; st->cnt = (int)C;
            st->pass = mm; st->done = done ? 1 : 0;
        }
        if (done) return;
    }

    matvec_tma_core<R, STAGES, false>(A, ld, nloc, n4, v, t + row0, ws, reinterpret_cast<float4*>(tma_smem), full_bar, empty_bar, red);
    const int nch = (n4 + MV_THREADS - 1) / MV_THREADS;
    const long long U = (long long)((nloc + R - 1) / R) * nch;
    if (tid == 0) s_last = (atomicAdd(ticket, 1u) == (unsigned int)(G - 1));
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    fold_split_bands<R>(t + row0, ws, U, G, nch, nloc, TMA_THREADS);
    if (tid == 0) {
        *ticket = 0u;
        st->matvecs = m + 1;
    }
}

// Fused IC pass on the upper triangle: P1 folds t_i from the partials of the previous launch's symmetric
// matvec on the fly, P2/P3 as in k_ic_pass, then the symmetric TMA main loop with z = v.
template <int STAGES, int OCC, bool FOLD>
__global__ void __launch_bounds__(SYM_THREADS, OCC)
k_ic_pass_sym_tma(const float* __restrict__ A, long long ld, int n, int n4, const uint8_t* __restrict__ keep, double* t, double* v,
                  double* B, double* rowpart, double* colpart, SymTab T, unsigned int* ticket, IcState* st, double* part) {
    namespace cg = cooperative_groups;
    if (st->done) return;
    cg::grid_group grid = cg::this_grid();
    extern __shared__ __align__(128) unsigned char tma_smem[];
    __shared__ __align__(8) uint64_t full_bar[STAGES];
    __shared__ __align__(8) uint64_t empty_bar[STAGES];
    __shared__ double sm[32];
    __shared__ double gred[2 * SYM_CW * SYM_R];
    __shared__ __align__(16) double cred[(SYM_NG - 1) * 1024];       // column partials of the groups g > 0 on their way to group 0
    const int tid = threadIdx.x;
    const int G = gridDim.x, bid = blockIdx.x;
    const int m = st->matvecs;
    const int gth = bid * SYM_THREADS + tid, nth = G * SYM_THREADS;

    if (m > 0) {
        const int mm = m - 1;
        const double cc_prev = st->cc, tol = st->tol;
        const int max_iter = st->max_iter;
        double s = 0.0, c = 0.0;
        if (FOLD) {
            // one rank: t_i is folded here by groups of 8 lanes (SYM_THREADS = 288 is a multiple of 8); lane 0 owns bin i
            for (int base = 0; base < n; base += nth / 8) {
                const int i = base + (gth >> 3);
                const double ti = sym_fold_group8(i, i < n, n4, rowpart, colpart, T);
                if (i < n && (tid & 7) == 0) {
                    t[i] = ti;
                    if (keep[i]) {
                        const double p = v[i] * ti;
                        s += p;
                        c += (p != 0.0) ? 1.0 : 0.0;
                    }
                }
            }
        } else {
            // sharded: k_sym_fold + an allreduce between the launches left the complete t
            for (int i = gth; i < n; i += nth) {
                if (keep[i]) {
                    const double p = v[i] * t[i];
                    s += p;
                    c += (p != 0.0) ? 1.0 : 0.0;
                }
            }
        }
        s = block_reduce<RED_SUM>(s, sm);
        c = block_reduce<RED_SUM>(c, sm);
        if (tid == 0) {
            part[bid] = s;
            part[G + bid] = c;
        }
        grid.sync();
        double a0 = 0.0, a1 = 0.0;
        for (int b = tid; b < G; b += SYM_THREADS) {
            a0 += __ldcg(part + b);
            a1 += __ldcg(part + G + b);
        }
        const double S = block_reduce<RED_SUM>(a0, sm), C = block_reduce<RED_SUM>(a1, sm);
        const double cc = (mm == 0) ? S : cc_prev;          // contact_count, lib/normalizeCore.pyx:36
        const double f = (mm == 0) ? 1.0 : sqrt(S / cc);    // :52
        const double g = (mm == 0) ? 1.0 : cc / S;          // :53
        const double mean = g * S / C;                      // :44
        double l1 = 0.0;
        for (int i = gth; i < n; i += nth) {
            if (keep[i]) {
                const double vi = v[i];
                const double d = g * (vi * t[i]);
                double bn = B[i];
                if (mm > 0) {
                    bn = f / vi;                            // :51-52
                    l1 += fabs(B[i] - bn);                  // :56
                    B[i] = bn;
                }
                v[i] = (d != 0.0) ? mean / (bn * d) : 1.0 / bn;     // :44-45
            }
        }
        l1 = block_reduce<RED_SUM>(l1, sm);
        if (tid == 0) part[2 * G + bid] = l1;
        grid.sync();
        double a2 = 0.0;
        for (int b = tid; b < G; b += SYM_THREADS) a2 += __ldcg(part + 2 * G + b);
        const double L1 = block_reduce<RED_SUM>(a2, sm);
        const bool done = (mm >= 2) && (L1 < tol || (mm - 1) > max_iter);       // :55-57
        if (bid == 0 && tid == 0) {
            st->cc = cc; st->S = S; st->l1 = L1; st->mean = mean; st->cnt = (int)C;
            st->pass = mm; st->done = done ? 1 : 0;
        }
        if (done) return;
    }
    // the previous launch's partials have been consumed by every CTA (grid barriers above, or m == 0)
    matvec_sym_core<SYM_NG, STAGES, false>(A, ld, n4, v, rowpart, colpart, T, reinterpret_cast<float4*>(tma_smem), full_bar, empty_bar, gred, cred);
    if (tid == 0) {
        __threadfence();
        if (atomicAdd(ticket, 1u) == (unsigned int)(G - 1)) {       // last CTA to issue its work: safe to bump the counter
            *ticket = 0u;
            st->matvecs = m + 1;
        }
    }
}

// ---------------------------------------------------------------------------------------------------
// Fused IC pass, deferred scalars (default on the upper-triangle path).  The bias update of lib/normalizeCore.pyx:42-53 is
//     v_next,i = mean / (B_next,i * d_i) = kappa / t_i          with ONE global scalar kappa = mean / (f g),
// because B_next,i = f / v_i and d_i = g v_i t_i (f = sqrt(S'/cc) :52, g = cc/S' :53, mean = g S'/C :44).  So the product of
// launch L runs on the UNSCALED vector u = 1/t' (t' = A u of the previous launch) the moment t' is folded -- the global sums
// (S', count, L1) are no longer between the fold and the product.  With v = alpha u:  t = alpha t',  S' = alpha^2 sum u t',
// alpha_next = kappa / alpha, and the bias of the pass closed by product m is B_i = (f_m / alpha_m) t'_{m-1,i} = beta_m t'_{m-1,i}
// on bins with data, Phi_m = f_1 ... f_m on kept bins whose row is empty (their v only ever multiplies zeros).  Launch L:
//   P0  producer warp issues the first tiles of A (they do not depend on u)
//   P1  per bin: fold t'_{L-1}; u_L = 1/t'; partial sum u_{L-1} t'_{L-1} and count; with beta/Phi of pass L-2 (known since launch
//       L-1): B_i, partial L1 of pass L-2                                     -> ONE grid barrier (u_L complete)
//   P2  every CTA folds the 3 x G partials in the same order: scalars of pass L-1, convergence of pass L-2 (:55-57) -> exit or
//   P3  the symmetric product on u_L.
// Convergence is seen one launch later than in k_ic_pass_sym_tma (one extra product per run), the vector phase shrinks from
// fold + 2 grid barriers + 3 block reductions + 2 sweeps to fold + 1 barrier.  Rounding differs from the literal order by a few
// ulp per pass (parity: bias 1e-9, equal pass counts -- tests/test_gpu_parity.py).  A bin whose t' changes between zero and
// non-zero from one pass to the next (impossible on a symmetric non-negative map) sets st->irregular; the host then reruns with
// the literal-order kernel.
//
// DIST: the map is row-sharded (circulant block assignment, see sym_ready); each rank folds its PARTIAL t' for all n bins and
// stores it straight into every peer's receive buffer over NVLink (peer-mapped memory, cudaIpc or same-process P2P); per-CTA flags
// replace a collective: CTA b of every rank owns the same bins, so CTA b only waits for the W flags of CTA b on the peers.  Every
// rank then adds the W partials in rank order -> bitwise identical t', u, scalars on all ranks; no NCCL call, no extra launch.
// ---------------------------------------------------------------------------------------------------
constexpr int XCHG_MAX_WORLD = 8;
struct IcXchg {
    int world, rank;
    unsigned int seq0;                          // flag value of launch 0 of this run (flags only ever grow)
    unsigned int* flags;                        // this rank's [world][G]: last launch for which CTA b of rank q delivered
    const double* recv;                         // this rank's [2][world][nv]: partial t' per source rank, double-buffered by launch parity
    unsigned int* peer_flags[XCHG_MAX_WORLD];   // the same two buffers on every rank, peer-mapped
    double* peer_recv[XCHG_MAX_WORLD];
};

__device__ __forceinline__ void st_release_sys_u32(unsigned int* p, unsigned int v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned int ld_acquire_sys_u32(const unsigned int* p) {
    unsigned int v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// Vector phase of the deferred-scalar pass for one launch: [DIST: this rank's partial t' of every bin -> every rank's receive
// buffer, per-CTA flags] fold, next vector u, partial sums.  LPB = lanes that share a bin.
template <int LPB, bool DIST>
__device__ __forceinline__ void symd_vector_phase(int n, int n4, long long nv, int L, const uint8_t* __restrict__ keep, double* tcur, const double* tm2,
                                                  const double* tm3, const double* uprev, double* ucur, double* B, const double* rowpart,
                                                  const double* colpart, const SymTab& T, IcState* st, const IcXchg& X, double beta, double phi,
                                                  long long* stamps, double& s, double& c, double& l1, int& irr) {
    constexpr int WPC = SYM_THREADS / 32, BPW = 32 / LPB;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int G = gridDim.x, bid = blockIdx.x;
    const int par = L & 1;
    if (DIST) {
        // ---- P1a: this rank's partial t' of every bin -> all ranks' receive buffers ----
        for (int q0 = 0; BPW * q0 < n; q0 += WPC * G) {
            const int i = BPW * (q0 + warp * G + bid) + lane / LPB;
            bool mine = false;
            const double pi = sym_fold_mlp<LPB>(i, i < n, n4, rowpart, colpart, T, &mine);
            // bins this rank never contributes to (columns of blocks it does not walk, outside its own rows) are not sent: the
            // peers' slots for them keep the zeros they were created with
            if (i < n && mine && (tid & (LPB - 1)) == 0) {
                for (int q = 0; q < X.world; ++q) X.peer_recv[q][((long long)par * X.world + X.rank) * nv + i] = pi;
            }
        }
        // The CTA barrier orders every thread's peer stores before the flag threads; their release store at system scope then
        // publishes all of them (release is cumulative over what happened before it through the barrier -- the pattern of
        // NCCL's post step and of grid.sync).  A system fence by all 256 threads here cost several microseconds per pass.
        __syncthreads();
        if (stamps != nullptr) stamps[12] = (long long)globaltimer_ns();
        if (tid < X.world) {
            const unsigned int want = X.seq0 + (unsigned int)L;
            __threadfence_system();
            st_release_sys_u32(X.peer_flags[tid] + (long long)X.rank * G + bid, want);
            const unsigned int* f = X.flags + (long long)tid * G + bid;
            const unsigned long long t0 = globaltimer_ns();
            while ((int)(ld_acquire_sys_u32(f) - want) < 0) {
                if (globaltimer_ns() - t0 > 20000000000ull) {       // 20 s: a peer died -- fail instead of hanging the box
                    st->error = 1;
                    break;
                }
            }
        }
        __syncthreads();
        if (stamps != nullptr) stamps[13] = (long long)globaltimer_ns();
    }
    // ---- P1: fold, next vector, partial sums ----
    for (int q0 = 0; BPW * q0 < n; q0 += WPC * G) {
        const int i = BPW * (q0 + warp * G + bid) + lane / LPB;
        double tp = 0.0;
        if (!DIST) tp = sym_fold_mlp<LPB>(i, i < n, n4, rowpart, colpart, T);
        if (i < n && (tid & (LPB - 1)) == 0) {
            if (DIST) {
                double r[XCHG_MAX_WORLD];
#pragma unroll
                for (int q = 0; q < XCHG_MAX_WORLD; ++q)                  // independent loads, then the sum in rank order
                    r[q] = q < X.world ? __ldcg(X.recv + ((long long)par * X.world + q) * nv + i) : 0.0;
#pragma unroll
                for (int q = 0; q < XCHG_MAX_WORLD; ++q) tp += r[q];
            }
            tcur[i] = tp;
            double un = 0.0;
            if (keep[i]) {
                const double p = uprev[i] * tp;
                s += p;
                c += (p != 0.0) ? 1.0 : 0.0;
                const bool ne = tp != 0.0;
                un = ne ? 1.0 / tp : 0.0;
                if (L >= 2 && ((tm2[i] != 0.0) != ne)) irr = 1;
                if (L >= 3) {
                    const double w = tm3[i];
                    const double bn = (w != 0.0) ? beta * w : phi;       // f / v (:51-52); empty kept bin: f_1 ... f_m
                    l1 += fabs(B[i] - bn);                               // :56
                    B[i] = bn;
                }
            }
            ucur[i] = un;
        }
    }
}

template <int STAGES, int OCC, bool DIST>
__global__ void __launch_bounds__(SYM_THREADS, OCC)
k_ic_pass_symd(const float* __restrict__ A, long long ld, int n, int n4, long long nv, const uint8_t* __restrict__ keep, double* tring,
               double* uring, double* B, double* rowpart, double* colpart, SymTab T, unsigned int* ticket, IcState* st, double* part,
               IcXchg X, int npass) {
    // npass: passes this launch may run.  1 = one cooperative launch per pass (the host enqueues them ahead of its convergence
    // check); large = the whole solve in ONE launch: the kernel boundary between passes becomes a grid barrier, and the first
    // tiles of the next pass are already in flight while the grid waits at it.
    namespace cg = cooperative_groups;
    if (st->done) return;
    cg::grid_group grid = cg::this_grid();
    extern __shared__ __align__(128) unsigned char tma_smem[];
    __shared__ __align__(8) uint64_t full_bar[STAGES];
    __shared__ __align__(8) uint64_t empty_bar[STAGES];
    __shared__ double sm[32];
    __shared__ double gred[2 * SYM_CW * SYM_R];
    __shared__ __align__(16) double cred[(SYM_NG - 1) * 1024];       // column partials of the groups g > 0 on their way to group 0
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int G = gridDim.x, bid = blockIdx.x;
    const int L0 = st->matvecs;                       // products completed so far; the partials hold product L0-1
    // vector phase: a warp takes 32 / LPB consecutive bins and consecutive warp-loads go to consecutive CTAs, so bins that are
    // expensive to fold (many column partials) are spread over the whole grid
    constexpr int WPC = SYM_THREADS / 32;
    const double tol = st->tol;
    const int max_iter = st->max_iter;
    float4* tiles = reinterpret_cast<float4*>(tma_smem);
    const long long su0 = T.cta_u0[bid], su1 = T.cta_u0[bid + 1];
  for (int it = 0; it < npass; ++it) {
    const int L = L0 + it;
    double* ucur = uring + (long long)(L & 1) * nv;

    long long* stamps = (T.dbg_stamps != nullptr && tid == 0) ? T.dbg_stamps + (long long)bid * 16 : nullptr;
    if (stamps != nullptr) {
        stamps[8] = (L == T.dbg_launch) ? 1 : 0;           // arms the two stamps taken inside matvec_sym_core
        if (L == T.dbg_launch) {
            stamps[0] = (long long)globaltimer_ns();
            unsigned int smid;
            asm volatile("mov.u32 %0, %smid;" : "=r"(smid));
            stamps[9] = smid;
            stamps[10] = T.cta_kf[bid];
            stamps[11] = T.cta_kl[bid];
        } else stamps = nullptr;
    }
    // ---- P0: barriers + the first tiles of this CTA's static share ----
    matvec_sym_range<SYM_NG, STAGES, false, SYM_BEGIN>(A, ld, n4, ucur, rowpart, colpart + (size_t)bid * T.nslots * SYM_NG * 1024, T, tiles, full_bar,
                                                       empty_bar, gred, cred, su0, su1, T.cta_kf[bid], T.cta_kl[bid], false);
    if (stamps != nullptr) stamps[1] = (long long)globaltimer_ns();
    if (it > 0) grid.sync();          // the previous pass's partials are complete (this replaces the kernel boundary)
    if (stamps != nullptr) stamps[14] = (long long)globaltimer_ns();
    const double alpha = st->alpha, beta = st->beta, phi = st->phi, cc_prev = st->cc;
    if (L > 0) {
        const double* uprev = uring + (long long)((L + 1) & 1) * nv;
        double* tcur = tring + (long long)((L + 2) % 3) * nv;            // t'_{L-1}
        const double* tm2 = tring + (long long)((L + 1) % 3) * nv;       // t'_{L-2}
        const double* tm3 = tring + (long long)(L % 3) * nv;             // t'_{L-3} = 1 / u_{L-2}
        double s = 0.0, c = 0.0, l1 = 0.0;
        int irr = 0;
        // lanes per bin: 2 while one round of the grid covers the map (16 bins per warp, short source chains), 1 beyond that
        // (32 bins per warp: half the rounds) -- the same choice on every rank (it depends on n only)
        if (n <= 16 * WPC * G)
            symd_vector_phase<2, DIST>(n, n4, nv, L, keep, tcur, tm2, tm3, uprev, ucur, B, rowpart, colpart, T, st, X, beta, phi, stamps, s, c, l1, irr);
        else
            symd_vector_phase<1, DIST>(n, n4, nv, L, keep, tcur, tm2, tm3, uprev, ucur, B, rowpart, colpart, T, st, X, beta, phi, stamps, s, c, l1, irr);
        if (irr) st->irregular = 1;
        if (stamps != nullptr) stamps[15] = (long long)globaltimer_ns();
        // one block reduction for the three sums (fixed tree)
        s = warp_sum(s); c = warp_sum(c); l1 = warp_sum(l1);
        if (lane == 0) { sm[warp] = s; gred[warp] = c; gred[16 + warp] = l1; }
        __syncthreads();
        if (tid == 0) {
            double a = 0.0, b = 0.0, d = 0.0;
#pragma unroll
            for (int w = 0; w < SYM_THREADS / 32; ++w) { a += sm[w]; b += gred[w]; d += gred[16 + w]; }
            part[bid] = a;
            part[G + bid] = b;
            part[2 * G + bid] = d;
        }
        if (stamps != nullptr) stamps[2] = (long long)globaltimer_ns();
        grid.sync();
        if (stamps != nullptr) stamps[3] = (long long)globaltimer_ns();
        // ---- P2: scalars of pass L-1, convergence of pass L-2 (every CTA, same order -> same bits) ----
        if (warp == 0) {
            double a0 = 0.0, a1 = 0.0, a2 = 0.0;
            for (int b = lane; b < G; b += 32) {
                a0 += __ldcg(part + b);
                a1 += __ldcg(part + G + b);
                a2 += __ldcg(part + 2 * G + b);
            }
            a0 = warp_sum(a0); a1 = warp_sum(a1); a2 = warp_sum(a2);
            if (lane == 0) { sm[0] = a0; sm[1] = a1; sm[2] = a2; }
        }
        __syncthreads();
        const double Sp = sm[0], C = sm[1], L1 = sm[2];
        const int mm = L - 1;                                 // the product these sums close
        const double S = alpha * alpha * Sp;                  // v't with v = alpha u, t = alpha t'
        const double cc = (mm == 0) ? S : cc_prev;            // contact_count, lib/normalizeCore.pyx:36
        const double f = (mm == 0) ? 1.0 : sqrt(S / cc);      // :52
        const double g = (mm == 0) ? 1.0 : cc / S;            // :53
        const double mean = g * S / C;                        // :44
        const double kappa = mean / (f * g);
        const int mp = L - 2;                                 // the pass whose L1 is complete now
        const bool done = (mp >= 2) && (L1 < tol || (mp - 1) > max_iter);       // :55-57
        const bool err = (DIST && st->error) || st->irregular;      // both written before the grid barrier
        if (bid == 0 && tid == 0) {
            st->cc = cc; st->S = S; st->mean = mean; st->cnt = (int)C;
            st->alpha = kappa / alpha; st->beta = f / alpha; st->phi = phi * f;
            if (mp >= 1) { st->l1 = L1; st->pass = mp; }
            st->done = (done || err) ? 1 : 0;
        }
        if (done || err) {
            // the tiles issued in P0 must land before the CTA may exit
            const long long npre = (su1 - su0 < (long long)STAGES) ? (su1 - su0) : (long long)STAGES;
            if (tid < (int)npre) mbar_wait(&full_bar[tid], 0u);
            __syncthreads();
            return;
        }
    }
    if (stamps != nullptr) stamps[4] = (long long)globaltimer_ns();
    matvec_sym_core<SYM_NG, STAGES, false, true>(A, ld, n4, ucur, rowpart, colpart, T, tiles, full_bar, empty_bar, gred, cred);
    if (tid == 0) {
        __threadfence();
        if (atomicAdd(ticket, 1u) == (unsigned int)(G - 1)) {       // last CTA: every CTA is past its chunk fetches
            *ticket = 0u;
            if (T.ndyn > 0) *T.dyn_counter = 0u;
            st->matvecs = L + 1;
            __threadfence();
        }
    }
  }
}

// s_i = keep_i ? 1/B_i : 0 -- the symmetric scale for the final apply
__global__ void k_ic_scale(int nv, int n, const uint8_t* __restrict__ keep, const double* __restrict__ B, double* s) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nv; i += gridDim.x * blockDim.x)
        s[i] = (i < n && keep[i]) ? 1.0 / B[i] : 0.0;
}

// ---------------------------------------------------------------------------------------------------
// KR vector step -- everything of lib/normalizeKnightRuiz.pyx:106-191, 271-286 except the matvec.
// One launch per matvec; `phase` says what the matvec just produced: 0 -> q = A x (init or outer Newton
// step), 1 -> q = A (x o p) (inner CG step).  The kernel advances the reference's control flow up to the
// next matvec and leaves its input in z.
// ---------------------------------------------------------------------------------------------------
struct KrState {
    double tol, delta, Delta, rt, stop_tol, g, etamax;
    double eta, rho_km1, rho_km2, rout, rold, innertol;
    int phase, init, k, outer, mvp, matvecs, done, pad;
};

struct KrVecs {
    double *x, *v, *rk, *y, *Z, *p, *w, *z;
};

__global__ void __launch_bounds__(EP_THREADS)
k_kr_init(KrState* st, int nv, const uint8_t* __restrict__ keep, KrVecs V, double tol, double delta, double Delta) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nv; i += gridDim.x * blockDim.x) {
        const double k = keep[i] ? 1.0 : 0.0;
        V.x[i] = k;       // x = e (:111, :118)
        V.z[i] = k;
        V.v[i] = 0; V.rk[i] = 0; V.y[i] = k; V.Z[i] = 0; V.p[i] = 0; V.w[i] = 0;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        st->tol = tol; st->delta = delta; st->Delta = Delta;
        st->rt = tol * tol;            // :119
        st->stop_tol = tol * 0.5;      // :116
        st->g = 0.9; st->etamax = 0.1; // :113-114
        st->eta = 0.1;                 // :115
        st->rho_km1 = 0; st->rho_km2 = 0; st->rout = 0; st->rold = 0; st->innertol = 0;
        st->phase = 0; st->init = 1; st->k = 0; st->outer = 0; st->mvp = 0; st->matvecs = 0; st->done = 0; st->pad = 0;
    }
}

__global__ void __cluster_dims__(EP_CL, 1, 1) __launch_bounds__(EP_THREADS)
k_kr_step(KrState* st, int n, const uint8_t* __restrict__ keep, const double* __restrict__ q, KrVecs V) {
    __shared__ double sm[32];
    __shared__ double slots[8];
    if (st->done) return;
    Team T(slots, sm);
    KrState s = *st;
    const int tid = T.gtid, nt = T.nthreads;
    bool to_outer = false;

    if (s.phase == 0) {
        // v = x.*(A*x); rk = 1 - v; rho_km1 = rk'*rk   (:120-122 / :169-171)
        double acc = 0.0;
        for (int i = tid; i < n; i += nt) {
            if (keep[i]) {
                const double vi = V.x[i] * q[i];
                const double r = 1.0 - vi;
                V.v[i] = vi;
                V.rk[i] = r;
                V.y[i] = 1.0;      // y = e for the next outer iteration (:278)
                acc += r * r;
            }
        }
        const double rho = T.reduce1<RED_SUM>(acc);
        s.rho_km1 = rho;
        if (s.init) {
            s.rout = rho; s.rold = rho; s.init = 0;          // :123-124
        } else {
            s.rout = rho;                                     // :172
            s.mvp += s.k + 1;                                 // :173
            const double rat = s.rout / s.rold;               // :176
            s.rold = s.rout;
            const double res_norm = sqrt(s.rout);
            const double eta_o = s.eta;
            s.eta = s.g * rat;                                // :180
            if (s.g * eta_o * eta_o > 0.1) s.eta = fmax(s.eta, s.g * eta_o * eta_o);        // :184-185
            s.eta = fmax(fmin(s.eta, s.etamax), s.stop_tol / res_norm);                     // :187
        }
        if (!(s.rout > s.rt)) {                               // :275
            s.done = 1;
            s.matvecs += 1;
            T.finish();                 // every CTA has read *st (all passed the reduction above)
            if (tid == 0) *st = s;
            return;
        }
        s.outer += 1;                                         // :276
        s.k = 0;
        s.innertol = fmax(s.eta * s.eta * s.rout, s.rt);      // :279
    } else {
        // w = x.*(A*(x.*p)) + v.*p ; alpha = rho_km1/(p'*w)   (:139-140)
        double acc = 0.0;
        for (int i = tid; i < n; i += nt) {
            if (keep[i]) {
                const double pi = V.p[i];
                const double wi = V.x[i] * q[i] + V.v[i] * pi;
                V.w[i] = wi;
                acc += pi * wi;
            }
        }
        const double pw = T.reduce1<RED_SUM>(acc);
        const double alpha = s.rho_km1 / pw;
        // ynew = y + ap ; distance to the cone boundary (:144-158)
        double mn = __longlong_as_double(0x7ff0000000000000LL), mx = __longlong_as_double(0xfff0000000000000LL);
        for (int i = tid; i < n; i += nt) {
            if (keep[i]) {
                const double yn = V.y[i] + alpha * V.p[i];
                mn = fmin(mn, yn);
                mx = fmax(mx, yn);
            }
        }
        mn = T.reduce1<RED_MIN>(mn);
        mx = T.reduce1<RED_MAX>(mx);
        if (mn <= s.delta) {
            if (s.delta != 0.0) {
                double gm = __longlong_as_double(0x7ff0000000000000LL);
                for (int i = tid; i < n; i += nt) {
                    if (keep[i]) {
                        const double ap = alpha * V.p[i];
                        if (ap < 0.0) gm = fmin(gm, (s.delta - V.y[i]) / ap);          // :150-151
                    }
                }
                const double gamma = T.reduce1<RED_MIN>(gm);
                for (int i = tid; i < n; i += nt)
                    if (keep[i]) V.y[i] = V.y[i] + gamma * (alpha * V.p[i]);           // :152
            }
            to_outer = true;
        } else if (mx >= s.Delta) {
            double gm = __longlong_as_double(0x7ff0000000000000LL);
            for (int i = tid; i < n; i += nt) {
                if (keep[i]) {
                    const double ap = alpha * V.p[i];
                    const double yn = V.y[i] + ap;
                    if (yn > s.Delta) gm = fmin(gm, (s.Delta - V.y[i]) / ap);          // :155-156
                }
            }
            const double gamma = T.reduce1<RED_MIN>(gm);
            for (int i = tid; i < n; i += nt)
                if (keep[i]) V.y[i] = V.y[i] + gamma * (alpha * V.p[i]);               // :157
            to_outer = true;
        } else {
            // y = ynew; rk = rk - alpha*w; Z = rk./v; rho_km1 = rk'*Z   (:159-163)
            double a2 = 0.0;
            for (int i = tid; i < n; i += nt) {
                if (keep[i]) {
                    V.y[i] = V.y[i] + alpha * V.p[i];
                    const double r = V.rk[i] - alpha * V.w[i];
                    const double Zi = r / V.v[i];
                    V.rk[i] = r;
                    V.Z[i] = Zi;
                    a2 += r * Zi;
                }
            }
            s.rho_km2 = s.rho_km1;
            s.rho_km1 = T.reduce1<RED_SUM>(a2);
        }
    }

    // ---- inner-loop test (:281) and preparation of the next matvec input ----
    if (!to_outer && s.rho_km1 > s.innertol) {
        s.k += 1;                                                                       // :127
        if (s.k == 1) {
            double acc = 0.0;
            for (int i = tid; i < n; i += nt) {
                if (keep[i]) {
                    const double Zi = V.rk[i] / V.v[i];                                 // :130
                    V.Z[i] = Zi;
                    V.p[i] = Zi;                                                        // :131
                    V.z[i] = V.x[i] * Zi;
                    acc += V.rk[i] * Zi;
                }
            }
            s.rho_km1 = T.reduce1<RED_SUM>(acc);                                 // :132
        } else {
            const double beta = s.rho_km1 / s.rho_km2;                                  // :134
            for (int i = tid; i < n; i += nt) {
                if (keep[i]) {
                    const double pn = V.Z[i] + beta * V.p[i];                           // :135
                    V.p[i] = pn;
                    V.z[i] = V.x[i] * pn;
                }
            }
        }
        s.phase = 1;
    } else {
        for (int i = tid; i < n; i += nt) {
            if (keep[i]) {
                const double xn = V.x[i] * V.y[i];                                      // :168
                V.x[i] = xn;
                V.z[i] = xn;
            }
        }
        s.phase = 0;
    }
    s.matvecs += 1;
    T.finish();
    if (tid == 0) *st = s;
}

// ---------------------------------------------------------------------------------------------------
// N3: transition probability matrix, lib/statDist.py:50-99.
//   kept row i: q_ij = fl32(A_ij / s_i), s_i the row sum (:83-84); minvalue = smallest non-zero q over the kept rows,
//   ALL columns (:86-87); output = q_ij where i and j are kept and q_ij != 0, minvalue everywhere else (:89-95).
// WRITE = false: the minimum pass (reads 4 nloc n bytes); WRITE = true: the output pass (reads + writes, 8 nloc n).
// Pad columns (j >= n) stay zero, like everywhere else.
// ---------------------------------------------------------------------------------------------------
template <int R, int OCC, bool WRITE>
__global__ void __launch_bounds__(MV_THREADS, OCC)
k_tp_pass(const float* __restrict__ A, float* __restrict__ Out, long long ld, int nloc, int n4, int n, int row0,
          const float* __restrict__ rsum, const uint8_t* __restrict__ keep, float fill, MinMax* mm) {
    const int tid = threadIdx.x;
    const int nch = (n4 + MV_THREADS - 1) / MV_THREADS;
    const int nb = (nloc + R - 1) / R;
    const long long U = (long long)nb * nch, G = gridDim.x, c = blockIdx.x;
    const long long u0 = unit_begin(c, U, G), u1 = unit_begin(c + 1, U, G);
    float mn = GCX_PINF, mx = GCX_NINF;
    const uchar4* __restrict__ k4 = reinterpret_cast<const uchar4*>(keep);
    for (long long u = u0; u < u1; ++u) {
        const int b = (int)(u / nch), k = (int)(u - (long long)b * nch);
        const int c4 = k * MV_THREADS + tid;
        if (c4 >= n4) continue;
        float4 a[R];
#pragma unroll
        for (int rr = 0; rr < R; ++rr) {
            const int r = b * R + rr;
            if (r < nloc) a[rr] = ld_stream_f4(reinterpret_cast<const float4*>(A + (long long)r * ld) + c4);
        }
        const uchar4 kc = __ldg(k4 + c4);
        const unsigned char kcol[4] = {kc.x, kc.y, kc.z, kc.w};
        const int j0 = 4 * c4;
#pragma unroll
        for (int rr = 0; rr < R; ++rr) {
            const int r = b * R + rr;
            if (r >= nloc) break;
            const int i = row0 + r;
            const bool ki = __ldg(keep + i) != 0;
            const float si = __ldg(rsum + i);
            const float in[4] = {a[rr].x, a[rr].y, a[rr].z, a[rr].w};
            if constexpr (!WRITE) {
                // minimum pass: for a positive row sum the float32 quotient is a non-decreasing function of a positive entry,
                // so the smallest non-zero quotient of the thread's four entries comes from their smallest positive one: one
                // division per row segment.  Negative entries, a non-positive row sum or an underflowing quotient take the
                // exact per-entry path.  (Pad columns are zero and never count.)
                if (!ki) continue;
                float rmin = GCX_PINF;
                bool odd = !(si > 0.f);
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    rmin = fminf(rmin, in[e] > 0.f ? in[e] : GCX_PINF);
                    odd |= in[e] < 0.f;
                }
                if (!odd && rmin != GCX_PINF) {
                    const float q = __fdiv_rn(rmin, si);
                    if (q != 0.f) mn = fminf(mn, q);
                    else odd = true;
                }
                if (odd) {
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        if (in[e] == 0.f) continue;
                        const float q = __fdiv_rn(in[e], si);
                        if (q != 0.f) mn = fminf(mn, q);
                    }
                }
            } else {
                float o[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    float q = 0.f;
                    if (ki && in[e] != 0.f) q = __fdiv_rn(in[e], si);
                    q = (kcol[e] && q != 0.f) ? q : fill;
                    if (j0 + e >= n) q = 0.f;
                    o[e] = q;
                    if (j0 + e < n) mm_update(q, mn, mx);
                }
                st_stream_f4(reinterpret_cast<float4*>(Out + (long long)r * ld) + c4, make_float4(o[0], o[1], o[2], o[3]));
            }
        }
    }
    mm_commit(mn, mx, mm);
}

// smallest non-zero entry of the kept block (lib/statDist.py:466-469 replaces the zeros of that block with it)
template <int R, int OCC>
__global__ void __launch_bounds__(MV_THREADS, OCC)
k_block_min(const float* __restrict__ A, long long ld, int nloc, int n4, int n, int row0, const uint8_t* __restrict__ keep, MinMax* mm) {
    const int tid = threadIdx.x;
    const int nch = (n4 + MV_THREADS - 1) / MV_THREADS;
    const int nb = (nloc + R - 1) / R;
    const long long U = (long long)nb * nch, G = gridDim.x, c = blockIdx.x;
    const long long u0 = unit_begin(c, U, G), u1 = unit_begin(c + 1, U, G);
    float mn = GCX_PINF, mx = GCX_NINF;
    const uchar4* __restrict__ k4 = reinterpret_cast<const uchar4*>(keep);
    for (long long u = u0; u < u1; ++u) {
        const int b = (int)(u / nch), k = (int)(u - (long long)b * nch);
        const int c4 = k * MV_THREADS + tid;
        if (c4 >= n4) continue;
        float4 a[R];
#pragma unroll
        for (int rr = 0; rr < R; ++rr) {
            const int r = b * R + rr;
            if (r < nloc) a[rr] = ld_stream_f4(reinterpret_cast<const float4*>(A + (long long)r * ld) + c4);
        }
        const uchar4 kc = __ldg(k4 + c4);
        const unsigned char kcol[4] = {kc.x, kc.y, kc.z, kc.w};
#pragma unroll
        for (int rr = 0; rr < R; ++rr) {
            const int r = b * R + rr;
            if (r >= nloc) break;
            if (!__ldg(keep + row0 + r)) continue;
            const float in[4] = {a[rr].x, a[rr].y, a[rr].z, a[rr].w};
#pragma unroll
            for (int e = 0; e < 4; ++e)
                if (kcol[e]) mm_update(in[e], mn, mx);          // keep is zero beyond n: pad columns never count
        }
    }
    mm_commit(mn, mx, mm);
}

// ---------------------------------------------------------------------------------------------------
// N3: stationary distribution, lib/statDist.py:436-498.  The reference takes the first eigenvector of a dense LAPACK
// eigendecomposition of the transposed kept block (O(n^3)); here the same vector -- the dominant left eigenvector of
// a positive matrix -- comes from the power iteration  x <- x P' / |x P'|_1 ,  P'_ij = P_ij or m0 where P_ij = 0.
// One matrix read (4 nloc n bytes) per iteration: the transposed product below.
//
// k_matvec_t: CTA (k, c) = 1024 columns x one chunk of `rpc` rows; every thread keeps 4 column sums in registers
// (fp64) over the chunk's rows and writes them to part[c][.]; k_sd_fold adds the chunks in order.  No-data rows have
// x_i = 0 and drop out; no-data columns are ignored by the fold.
// ---------------------------------------------------------------------------------------------------
constexpr int MVT_UNR = 8;

__global__ void __launch_bounds__(MV_THREADS)
k_matvec_t(const float* __restrict__ A, long long ld, int nloc, int n4, int row0, int rpc, const double* __restrict__ x, float m0,
           double* __restrict__ part, const int* __restrict__ done) {
    if (done && *done) return;
    const int c4 = blockIdx.x * MV_THREADS + threadIdx.x;
    if (c4 >= n4) return;
    const int r0 = blockIdx.y * rpc, r1 = min(r0 + rpc, nloc);
    double acc[4] = {0.0, 0.0, 0.0, 0.0};
    const float4* __restrict__ base = reinterpret_cast<const float4*>(A) + c4;
    const long long ld4 = ld / 4;
    auto take = [&](const float4& a, double xv) {
        acc[0] = fma(xv, (double)(a.x == 0.f ? m0 : a.x), acc[0]);
        acc[1] = fma(xv, (double)(a.y == 0.f ? m0 : a.y), acc[1]);
        acc[2] = fma(xv, (double)(a.z == 0.f ? m0 : a.z), acc[2]);
        acc[3] = fma(xv, (double)(a.w == 0.f ? m0 : a.w), acc[3]);
    };
    int r = r0;
    for (; r + MVT_UNR <= r1; r += MVT_UNR) {
        float4 a[MVT_UNR];
        double xv[MVT_UNR];
#pragma unroll
        for (int u = 0; u < MVT_UNR; ++u) {
            a[u] = ld_stream_f4(base + (long long)(r + u) * ld4);
            xv[u] = __ldg(x + row0 + r + u);
        }
#pragma unroll
        for (int u = 0; u < MVT_UNR; ++u) take(a[u], xv[u]);
    }
    for (; r < r1; ++r) take(ld_stream_f4(base + (long long)r * ld4), __ldg(x + row0 + r));
    double2* dst = reinterpret_cast<double2*>(part + (long long)blockIdx.y * ld + 4 * (long long)c4);
    dst[0] = make_double2(acc[0], acc[1]);
    dst[1] = make_double2(acc[2], acc[3]);
}

struct SdState {
    double tol, delta, prev, sum;
    int iters, max_iter, done, matvecs, kept, pad;
};

__global__ void __cluster_dims__(EP_CL, 1, 1) __launch_bounds__(EP_THREADS)
k_sd_init(SdState* st, int nv, int n, const uint8_t* __restrict__ keep, double* x, double tol, int max_iter) {
    __shared__ double sm[32];
    __shared__ double slots[8];
    Team T(slots, sm);
    double cnt = 0.0;
    for (int i = T.gtid; i < n; i += T.nthreads) cnt += keep[i] ? 1.0 : 0.0;
    const double K = T.reduce1<RED_SUM>(cnt);
    for (int i = T.gtid; i < nv; i += T.nthreads) x[i] = (i < n && keep[i]) ? 1.0 / K : 0.0;
    if (T.gtid == 0) {
        st->tol = tol; st->delta = 0.0; st->prev = __longlong_as_double(0x7ff0000000000000LL); st->sum = 0.0;
        st->iters = 0; st->max_iter = max_iter; st->done = 0; st->matvecs = 0; st->kept = (int)K; st->pad = 0;
    }
    T.finish();
}

__global__ void __launch_bounds__(256)
k_sd_fold(const SdState* __restrict__ st, int nv, int n, int nrc, long long ld, const uint8_t* __restrict__ keep,
          const double* __restrict__ part, double* __restrict__ y) {
    if (st->done) return;
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < nv; j += gridDim.x * blockDim.x) {
        double s = 0.0;
        if (j < n && keep[j])
            for (int c = 0; c < nrc; ++c) s += part[(long long)c * ld + j];
        y[j] = s;
    }
}

// x <- y / sum(y); L1 change; stop below tol, at the iteration cap, or when the change stalls at the rounding floor
__global__ void __cluster_dims__(EP_CL, 1, 1) __launch_bounds__(EP_THREADS)
k_sd_step(SdState* st, int n, const uint8_t* __restrict__ keep, const double* __restrict__ y, double* x) {
    __shared__ double sm[32];
    __shared__ double slots[8];
    if (st->done) return;
    Team T(slots, sm);
    const SdState s = *st;
    double acc = 0.0;
    for (int i = T.gtid; i < n; i += T.nthreads)
        if (keep[i]) acc += y[i];
    const double S = T.reduce1<RED_SUM>(acc);
    double d = 0.0;
    for (int i = T.gtid; i < n; i += T.nthreads) {
        if (keep[i]) {
            const double yn = y[i] / S;
            d += fabs(yn - x[i]);
            x[i] = yn;
        }
    }
    const double delta = T.reduce1<RED_SUM>(d);
    if (T.gtid == 0) {
        const int it = s.iters + 1;
        st->sum = S; st->delta = delta; st->prev = delta; st->iters = it; st->matvecs = s.matvecs + 1;
        st->done = (delta < s.tol || it >= s.max_iter || (delta >= s.prev && delta < 1e-11)) ? 1 : 0;
    }
    T.finish();
}

// TMA-bulk variant of the transposed product (the default): persistent grid, units = (strip g of KG column chunks,
// row r) in g-major order split evenly over the CTAs, so a CTA walks down the rows of one KG x 4 KB wide column strip (two
// strips at most when its share crosses a strip boundary -> `slots` partial rows per CTA).  One producer lane keeps
// STAGES row segments (one contiguous KG x 4 KB copy each) in flight behind mbarriers; the 8 consumer warps add
// x_i * P'_ij into 4 KG register column sums per thread -- no cross-thread reduction at all.
// part: [grid][slots][KG * 1024] doubles; k_sd_fold_t adds the CTAs of a strip in CTA order.
struct MvtTab {
    const int* cta_gf;      // [G] first strip of every CTA
    const int* strip_lo;    // [nstrips] first / last CTA holding units of the strip
    const int* strip_hi;
    int slots;
};

template <int KG, int STAGES, int OCC>
__global__ void __launch_bounds__(TMA_THREADS, OCC)
k_matvec_t_tma(const float* __restrict__ A, long long ld, int nloc, int n4, int row0, const double* __restrict__ x, float m0,
               double* __restrict__ part, MvtTab T, const int* __restrict__ done) {
    if (done != nullptr && *done) return;
    extern __shared__ __align__(128) unsigned char tma_smem[];
    __shared__ __align__(8) uint64_t full_bar[STAGES];
    __shared__ __align__(8) uint64_t empty_bar[STAGES];
    float4* tiles = reinterpret_cast<float4*>(tma_smem);
    constexpr int SW4 = KG * MV_THREADS;                  // float4 columns per strip
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nst = (n4 + SW4 - 1) / SW4;
    const long long U = (long long)nst * nloc, G = gridDim.x, c = blockIdx.x;
    const long long u0 = unit_begin(c, U, G), u1 = unit_begin(c + 1, U, G);
    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], MV_THREADS / 32);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (u1 <= u0) return;
    int g = (int)(u0 / nloc), r = (int)(u0 - (long long)g * nloc);      // the only 64-bit divisions of the kernel
    const int gfirst = g;
    if (warp == MV_THREADS / 32) {
        if (lane == 0) {
            int stage = 0;
            unsigned int phase = 0;
            for (long long u = u0; u < u1; ++u) {
                mbar_wait(&empty_bar[stage], phase ^ 1u);
                const int c0 = g * SW4;
                const unsigned int bytes = (unsigned int)min(SW4, n4 - c0) * 16u;
                mbar_expect_tx(&full_bar[stage], bytes);
                bulk_g2s(tiles + (size_t)stage * SW4, A + (long long)r * ld + 4LL * c0, bytes, &full_bar[stage]);
                if (++stage == STAGES) {
                    stage = 0;
                    phase ^= 1u;
                }
                if (++r == nloc) {
                    r = 0;
                    ++g;
                }
            }
        }
        return;
    }
    double acc[KG][4];
#pragma unroll
    for (int kk = 0; kk < KG; ++kk) acc[kk][0] = acc[kk][1] = acc[kk][2] = acc[kk][3] = 0.0;
    auto flush = [&](int gg) {
        double* base = part + ((long long)c * T.slots + (gg - gfirst)) * (4LL * SW4);
#pragma unroll
        for (int kk = 0; kk < KG; ++kk) {
            double2* dst = reinterpret_cast<double2*>(base + 4 * (kk * MV_THREADS + tid));
            dst[0] = make_double2(acc[kk][0], acc[kk][1]);
            dst[1] = make_double2(acc[kk][2], acc[kk][3]);
            acc[kk][0] = acc[kk][1] = acc[kk][2] = acc[kk][3] = 0.0;
        }
    };
    int stage = 0;
    unsigned int phase = 0;
    for (long long u = u0; u < u1; ++u) {
        const double xv = __ldg(x + row0 + r);
        const int c0 = g * SW4;
        mbar_wait(&full_bar[stage], phase);
        if (xv != 0.0) {                                       // no-data rows contribute nothing
#pragma unroll
            for (int kk = 0; kk < KG; ++kk) {
                if (c0 + kk * MV_THREADS + tid < n4) {
                    const float4 a = tiles[(size_t)stage * SW4 + kk * MV_THREADS + tid];
                    acc[kk][0] = fma(xv, (double)(a.x == 0.f ? m0 : a.x), acc[kk][0]);
                    acc[kk][1] = fma(xv, (double)(a.y == 0.f ? m0 : a.y), acc[kk][1]);
                    acc[kk][2] = fma(xv, (double)(a.z == 0.f ? m0 : a.z), acc[kk][2]);
                    acc[kk][3] = fma(xv, (double)(a.w == 0.f ? m0 : a.w), acc[kk][3]);
                }
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty_bar[stage]);
        if (++stage == STAGES) {
            stage = 0;
            phase ^= 1u;
        }
        if (++r == nloc) {
            flush(g);
            r = 0;
            ++g;
        }
    }
    if (r != 0) flush(g);
}

template <int KG>
__global__ void __launch_bounds__(256)
k_sd_fold_t(const SdState* __restrict__ st, int nv, int n, const uint8_t* __restrict__ keep, const double* __restrict__ part, MvtTab T,
            double* __restrict__ y) {
    if (st->done) return;
    constexpr int SW = KG * 1024;
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < nv; j += gridDim.x * blockDim.x) {
        double s = 0.0;
        if (j < n && keep[j]) {
            const int g = j / SW, o = j - g * SW;
            const int lo = T.strip_lo[g], hi = T.strip_hi[g];
            for (int cta = lo; cta <= hi; ++cta) s += part[((long long)cta * T.slots + (g - T.cta_gf[cta])) * SW + o];
        }
        y[j] = s;
    }
}

// ---------------------------------------------------------------------------------------------------
// N4: pairwise-masked Pearson correlation of the rows of X = (kept block)^T, lib/corrMatrixCoreSRC.c.
// For a pair (i, j) the reference sums over the entries where BOTH rows differ from the mask value (:146-152, :181-186):
//   cov_ij = sum (x - xbar)(y - ybar) / N,  xbar = sum x / N  over those N entries, 0 unless N > 3 (:171);
//   corr_ij = cov_ij / sqrt(var_i var_j),   var_i = cov(x_i, x_i) with x_i's OWN mask only (:115-116) -- so |corr| can exceed 1.
// With X0 = x where unmasked else 0 and I = the 0/1 indicator, the pair sums are three GEMM-shaped products,
//   N = I I^T,  Sx = X0 I^T (sum of x_i over the common entries; sum of y is its transpose),  Sxy = X0 X0^T,
// and cov_ij = (Sxy_ij - Sx_ij Sx_ji / N_ij) / N_ij.  The products are plain library GEMMs (cuBLAS, nothing to fuse): the counts N
// from the bf16 copy of I with float32 accumulation on the tensor cores (0/1 inputs, integer sums below 2^24: exact), Sx as a
// float64 GEMM, the symmetric Sxy as a float64 SYRK (half the work; only the triangle the epilogue reads).  Everything around
// them is below.  var_i repeats the reference's sequential two-pass arithmetic
// (no FMA contraction), so the `var == 0` tests of :121 decide exactly as the reference does.
// ---------------------------------------------------------------------------------------------------
struct CorrPrep {
    int has_mask, logspace, has_vmin, has_vmax;
    double mask;
    float vmin, vmax, maskf;
};

__device__ __forceinline__ void corr_value(float a, const CorrPrep& P, double& x0, double& ind) {
    if (P.has_vmin && a <= P.vmin) a = P.maskf;                   // lib/corrMatrix.py:96-99 (on the float32 map)
    if (P.has_vmax && a >= P.vmax) a = P.maskf;
    double v = (double)a;                                          // :109
    if (P.logspace) {                                              // :111-113
        v = log10(v);
        if (isinf(v)) v = 0.0;
    }
    const bool on = !P.has_mask || v != P.mask;
    ind = on ? 1.0 : 0.0;
    x0 = on ? v : 0.0;
}

// X[i][k] = f(A[kept[k]][kept[i]]) -- compaction + transpose (:108-109) through a 32 x 32 shared tile; writes X0 and I
__global__ void __launch_bounds__(256)
k_corr_prep(const float* __restrict__ A, long long ld, const int* __restrict__ kept, int m, CorrPrep P, double* __restrict__ X0,
            double* __restrict__ I, unsigned short* __restrict__ Ib, int mb) {
    __shared__ float tile[32][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;          // 32 x 8
    const int i0 = blockIdx.x * 32, k0 = blockIdx.y * 32;
    // read: rows kept[k0 + r], columns kept[i0 + tx] (consecutive kept columns are close in memory)
    const int ci = i0 + tx < m ? kept[i0 + tx] : -1;
    for (int r = ty; r < 32; r += 8) {
        const int k = k0 + r;
        tile[r][tx] = (k < m && ci >= 0) ? A[(long long)kept[k] * ld + ci] : 0.f;
    }
    __syncthreads();
    for (int r = ty; r < 32; r += 8) {
        const int i = i0 + r, k = k0 + tx;
        if (i < m && k < m) {
            double x0, ind;
            corr_value(tile[tx][r], P, x0, ind);
            X0[(long long)i * m + k] = x0;
            I[(long long)i * m + k] = ind;
            Ib[(long long)i * mb + k] = ind != 0.0 ? (unsigned short)0x3F80 : (unsigned short)0;     // bf16 1.0 / 0.0, row pitch mb
        }
    }
}

// host float64 input (inner seam, lib/_corrMatrixCore.pyx:49): split into X0 and I in place (X holds the raw rows)
__global__ void __launch_bounds__(256)
k_corr_prep_f64(double* __restrict__ X, double* __restrict__ I, unsigned short* __restrict__ Ib, int m, int mb, int has_mask, double mask) {
    for (int i = blockIdx.x; i < m; i += gridDim.x) {
        for (int k = threadIdx.x; k < m; k += blockDim.x) {
            const long long t = (long long)i * m + k;
            const double v = X[t];
            const bool on = !has_mask || v != mask;
            I[t] = on ? 1.0 : 0.0;
            Ib[(long long)i * mb + k] = on ? (unsigned short)0x3F80 : (unsigned short)0;
            X[t] = on ? v : 0.0;
        }
    }
}

// var_i = cov(x_i, x_i) in the reference's own order and operations (:140-192): sequential sum, mean, sequential sum of squared
// deviations, no fused multiply-add.  One thread per row.
__global__ void __launch_bounds__(128)
k_corr_selfvar(const double* __restrict__ X0, const double* __restrict__ I, int m, double* __restrict__ var) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    const double* x = X0 + (long long)i * m;
    const double* w = I + (long long)i * m;
    double xsum = 0.0, cnt = 0.0;
    for (int k = 0; k < m; ++k) {
        if (w[k] != 0.0) {
            xsum = __dadd_rn(xsum, x[k]);
            cnt += 1.0;
        }
    }
    double v = 0.0;
    if (cnt > 3.0) {
        const double mean = xsum / cnt;
        double dev = 0.0;
        for (int k = 0; k < m; ++k) {
            if (w[k] != 0.0) {
                const double d = __dsub_rn(x[k], mean);
                dev = __dadd_rn(dev, __dmul_rn(d, d));
            }
        }
        v = dev / cnt;
    }
    var[i] = v;
}

// The same arithmetic -- every row's sums are still ONE sequential chain of non-fused float64 operations in index order, so the
// variances and the var == 0 decisions stay bit-identical -- with coalesced traffic: a warp owns 32 rows, cp.async stages 32 x 32
// tiles of X0 and I (row segments of 256 contiguous bytes) through a double-buffered, padded shared-memory tile while the lanes
// (one row each) walk the previous tile.  k_corr_selfvar read 8 useful bytes of every 32-byte sector (9.7 ms at m = 24,684).
constexpr int SV_WARPS = 4, SV_STAGES = 2, SV_PITCH = 33;
__device__ __forceinline__ void cp_async8(void* smem_dst, const void* gsrc) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_u32(smem_dst)), "l"(gsrc) : "memory");
}
__global__ void __launch_bounds__(SV_WARPS * 32)
k_corr_selfvar_tiled(const double* __restrict__ X0, const double* __restrict__ I, int m, double* __restrict__ var) {
    extern __shared__ double sv_smem[];        // [warp][stage][x | w][32][SV_PITCH]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int row0 = (blockIdx.x * SV_WARPS + warp) * 32;
    if (row0 >= m) return;
    double* base = sv_smem + (size_t)warp * SV_STAGES * 2 * 32 * SV_PITCH;
    const int nt = (m + 31) / 32;
    const int myrow = row0 + lane;
    auto issue = [&](int t, int stage) {
        double* sx = base + (size_t)stage * 2 * 32 * SV_PITCH;
        double* sw = sx + 32 * SV_PITCH;
        const int k = t * 32 + lane;
        if (k < m) {
#pragma unroll 8
            for (int rr = 0; rr < 32; ++rr) {
                const int r = row0 + rr;
                if (r < m) {
                    cp_async8(sx + rr * SV_PITCH + lane, X0 + (long long)r * m + k);
                    cp_async8(sw + rr * SV_PITCH + lane, I + (long long)r * m + k);
                }
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    double xsum = 0.0, cnt = 0.0, mean = 0.0, dev = 0.0;
    for (int pass = 0; pass < 2; ++pass) {
        issue(0, 0);
        for (int t = 0; t < nt; ++t) {
            if (t + 1 < nt) issue(t + 1, (t + 1) & 1);
            else asm volatile("cp.async.commit_group;" ::: "memory");
            asm volatile("cp.async.wait_group 1;" ::: "memory");
            __syncwarp();
            const double* sx = base + (size_t)(t & 1) * 2 * 32 * SV_PITCH + lane * SV_PITCH;
            const double* sw = sx + 32 * SV_PITCH;
            const int kmax = min(32, m - t * 32);
            if (myrow < m) {
                if (pass == 0) {
                    for (int k = 0; k < kmax; ++k) {
                        if (sw[k] != 0.0) {
                            xsum = __dadd_rn(xsum, sx[k]);
                            cnt += 1.0;
                        }
                    }
                } else if (cnt > 3.0) {
                    for (int k = 0; k < kmax; ++k) {
                        if (sw[k] != 0.0) {
                            const double d = __dsub_rn(sx[k], mean);
                            dev = __dadd_rn(dev, __dmul_rn(d, d));
                        }
                    }
                }
            }
            __syncwarp();          // the tile is free before the next prefetch overwrites it
        }
        if (pass == 0) mean = cnt > 3.0 ? xsum / cnt : 0.0;
    }
    if (myrow < m) var[myrow] = cnt > 3.0 ? dev / cnt : 0.0;
}

// tile (ti, tj), tj <= ti: corr for the pairs j < i, written to (i, j) and mirrored to (j, i) like :36-37; the diagonal stays 0.
// OUT32: scatter into the full float32 map at the kept indices; else dense float64 m x m.
template <bool OUT32>
__global__ void __launch_bounds__(256)
k_corr_finish(const double* __restrict__ N, const float* __restrict__ Nf, const double* __restrict__ Sx, const double* __restrict__ Sxy,
              int nf_pitch, int nf_int, const double* __restrict__ var,
              int m, const int* __restrict__ kept, float* __restrict__ out32, long long ld, double* __restrict__ out64, int cov_mode) {
    const int ti = blockIdx.x, tj = blockIdx.y;
    if (tj > ti) return;
    __shared__ double sT[32][33];                                    // Sx tile (tj, ti): holds Sx[j][i]
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    for (int r = ty; r < 32; r += 8) {
        const int j = tj * 32 + r, i = ti * 32 + tx;
        sT[r][tx] = (i < m && j < m) ? Sx[(long long)j * m + i] : 0.0;
    }
    __syncthreads();
    for (int r = ty; r < 32; r += 8) {
        const int i = ti * 32 + r, j = tj * 32 + tx;
        if (i >= m || j >= m || j > i || (j == i && !cov_mode)) continue;
        const long long o = (long long)i * m + j;
        // float32 counts (bf16 product) are exact: integers below 2^24; the sliced routes leave int32 counts in the same buffer
        const double cnt = !Nf ? N[o] : (nf_int ? (double)reinterpret_cast<const int*>(Nf)[(long long)i * nf_pitch + j] : (double)Nf[(long long)i * nf_pitch + j]);
        double cov = 0.0;
        if (cnt > 3.0) cov = (Sxy[o] - Sx[o] * sT[tx][r] / cnt) / cnt;
        const double vi = var[i], vj = var[j];
        double c;
        if (cov_mode) {
            c = (i == j) ? vi : cov;                                               // _covarianceMatrix, :83-92 (j <= i)
        } else {
            c = (vi == 0.0 || vj == 0.0) ? 0.0 : cov / sqrt(vi * vj);              // :121-128
            if (c != c) c = 0.0;                                                   // lib/corrMatrix.py:120
        }
        if (OUT32) {
            const int gi = kept[i], gj = kept[j];
            out32[(long long)gi * ld + gj] = (float)c;
            out32[(long long)gj * ld + gi] = (float)c;
        } else {
            out64[o] = c;
            out64[(long long)j * m + i] = c;
        }
    }
}

// Integer-valued maps (raw counts): X0 = sum_t 128^t s_t with 7-bit slices s_t in [0, 127] stored as int8.  A product of two
// slices (or of a slice and the 0/1 indicator) sums to at most 127^2 m < 2^31, so the int8 tensor-core GEMM with int32 accumulation is
// EXACT, and  Sx = sum_t 128^t (S_t I^T),  Sxy = sum_{a,b} 128^(a+b) (S_a S_b^T)  replace the float64 GEMM and SYRK.
// k_corr_intcheck: flag[0] |= 1 when an entry is not an integer in [0, 2^28); flag[1] = max entry (-> number of slices).
__global__ void __launch_bounds__(256)
k_corr_intcheck(const double* __restrict__ X0, long long total, unsigned int* flag) {
    int bad = 0;
    unsigned int mx = 0;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
        const double v = X0[t];
        const bool ok = v >= 0.0 && v < 268435456.0 && v == floor(v);
        bad |= !ok;
        if (ok) mx = max(mx, (unsigned int)v);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if (__syncthreads_or(bad) && threadIdx.x == 0) atomicOr(flag, 1u);
    if ((threadIdx.x & 31) == 0 && mx) atomicMax(flag + 1, mx);
}

// all `ns` slices of X0 and the int8 indicator, row pitch mb (pads stay 0)
__global__ void __launch_bounds__(256)
k_corr_slices8(const double* __restrict__ X0, const double* __restrict__ I, int m, int mb, int ns, signed char* __restrict__ S8,
               signed char* __restrict__ I8) {
    const long long plane = (long long)mb * mb;
    for (int i = blockIdx.x; i < m; i += gridDim.x) {
        const double* x = X0 + (long long)i * m;
        const double* w = I + (long long)i * m;
        for (int k = threadIdx.x; k < m; k += blockDim.x) {
            const unsigned int v = (unsigned int)x[k];
            const long long o = (long long)i * mb + k;
            for (int t = 0; t < ns; ++t) S8[t * plane + o] = (signed char)((v >> (7 * t)) & 127u);
            I8[o] = w[k] != 0.0 ? 1 : 0;
        }
    }
}

// dst = (first ? 0 : dst) + scale * T   (T: int32, row pitch mb; dst: float64 m x m)
__global__ void __launch_bounds__(256)
k_corr_accum32(const int* __restrict__ T, int m, int mb, double scale, int first, double* __restrict__ dst) {
    for (int i = blockIdx.x; i < m; i += gridDim.x) {
        const int* src = T + (long long)i * mb;
        double* d = dst + (long long)i * m;
        for (int k = threadIdx.x; k < m; k += blockDim.x) {
            const double v = scale * (double)src[k];
            d[k] = first ? v : d[k] + v;
        }
    }
}

// Non-integer maps whose entries are float32 numbers (everything but log space): Sx = X0 I^T stays EXACT on the int8 tensor cores.
// Every entry of row i is an integer multiple of 2^(lo_i) (lo = smallest exponent - 24) and below 2^(hi_i), so
// q = |x| 2^(7K - hi_i) is an integer below 2^(7K) as soon as 7K >= hi_i - lo_i; its K seven-bit digits (sign carried along) are the
// slices and  Sx_ij = 2^(hi_i) sum_t 2^(-7(t+1)) (S_t I^T)_ij  with exact int32 products -- K <= 9 products of ~7 ms instead of the
// 832 ms float64 GEMM.  k_corr_f32check finds hi_i, the largest hi_i - lo_i (info[1]) and raises info[0] when an entry is not a
// float32 number; the caller falls back to the float64 GEMM then, or when more than 63 bits would be needed.
__global__ void __launch_bounds__(256)
k_corr_f32check(const double* __restrict__ X0, int m, int* __restrict__ rhi, unsigned int* info) {
    __shared__ int s_hi[8], s_lo[8], s_bad[8];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int i = blockIdx.x; i < m; i += gridDim.x) {
        const double* x = X0 + (long long)i * m;
        int hi = -100000, lo = 100000, bad = 0;
        for (int k = threadIdx.x; k < m; k += blockDim.x) {
            const double v = x[k];
            if (v == 0.0) continue;
            bad |= !((double)(float)v == v) || !isfinite(v);
            int e;
            frexp(v, &e);                       // |v| = f 2^e, f in [0.5, 1)
            hi = max(hi, e);
            lo = min(lo, e - 24);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, o));
            lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, o));
            bad |= __shfl_xor_sync(0xffffffffu, bad, o);
        }
        if (lane == 0) { s_hi[warp] = hi; s_lo[warp] = lo; s_bad[warp] = bad; }
        __syncthreads();
        if (threadIdx.x == 0) {
            for (int w = 1; w < 8; ++w) { hi = max(hi, s_hi[w]); lo = min(lo, s_lo[w]); bad |= s_bad[w]; }
            const bool empty = hi < lo;
            rhi[i] = empty ? 0 : hi;
            if (bad) atomicOr(info, 1u);
            if (!empty) atomicMax(info + 1, (unsigned int)(hi - lo));
        }
        __syncthreads();
    }
}

__global__ void __launch_bounds__(256)
k_corr_fslices8(const double* __restrict__ X0, const double* __restrict__ I, int m, int mb, int K, const int* __restrict__ rhi,
                signed char* __restrict__ S8, signed char* __restrict__ I8) {
    const long long plane = (long long)mb * mb;
    for (int i = blockIdx.x; i < m; i += gridDim.x) {
        const double* x = X0 + (long long)i * m;
        const double* w = I + (long long)i * m;
        const int sh = 7 * K - rhi[i];
        for (int k = threadIdx.x; k < m; k += blockDim.x) {
            const double v = x[k];
            const unsigned long long q = (unsigned long long)ldexp(fabs(v), sh);        // exact: an integer below 2^(7K)
            const long long o = (long long)i * mb + k;
            for (int t = 0; t < K; ++t) {
                const int sl = (int)((q >> (7 * (K - 1 - t))) & 127ull);
                S8[t * plane + o] = (signed char)(v < 0.0 ? -sl : sl);
            }
            I8[o] = w[k] != 0.0 ? 1 : 0;
        }
    }
}

// Sx[i][j] (+)= T[i][j] 2^(hi_i - shift)
__global__ void __launch_bounds__(256)
k_corr_faccum(const int* __restrict__ T, int m, int mb, const int* __restrict__ rhi, int shift, int first, double* __restrict__ Sx) {
    for (int i = blockIdx.x; i < m; i += gridDim.x) {
        const int* src = T + (long long)i * mb;
        double* d = Sx + (long long)i * m;
        const int e = rhi[i] - shift;
        for (int k = threadIdx.x; k < m; k += blockDim.x) {
            const double v = ldexp((double)src[k], e);
            d[k] = first ? v : d[k] + v;
        }
    }
}

// Sxy the same way while it beats the SYRK (K <= 8: at most 36 products):  Sxy_ij = 2^(hi_i + hi_j) sum_{a,b} 2^(-7(a+b+2)) (S_a S_b^T)_ij,
// ALL pairs (nothing truncated: exact int32 products, exact scalings, float64 roundings only in the final sum).  S_b S_a^T is the
// transpose of S_a S_b^T, so only a <= b is multiplied and this kernel adds T + T^T on the lower triangle -- the only part the
// epilogue reads:  Sxy[i][j] (+)= (T[i][j] + (sym ? T[j][i] : 0)) 2^(hi_i + hi_j - shift),  j <= i.
__global__ void __launch_bounds__(256)
k_corr_faccum_sxy(const int* __restrict__ T, int m, int mb, const int* __restrict__ rhi, int shift, int sym, int first,
                  double* __restrict__ Sxy) {
    const int ti = blockIdx.x, tj = blockIdx.y;
    if (tj > ti) return;
    __shared__ int sT[32][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    if (sym) {
        for (int r = ty; r < 32; r += 8) {
            const int j = tj * 32 + r, i = ti * 32 + tx;
            sT[r][tx] = (i < m && j < m) ? T[(long long)j * mb + i] : 0;
        }
        __syncthreads();
    }
    for (int r = ty; r < 32; r += 8) {
        const int i = ti * 32 + r, j = tj * 32 + tx;
        if (i >= m || j >= m || j > i) continue;
        double v = (double)T[(long long)i * mb + j];
        if (sym) v += (double)sT[tx][r];
        v = ldexp(v, rhi[i] + rhi[j] - shift);
        const long long o = (long long)i * m + j;
        Sxy[o] = first ? v : Sxy[o] + v;
    }
}

}  // namespace gcx
