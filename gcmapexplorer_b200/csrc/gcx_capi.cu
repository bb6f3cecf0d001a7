// gcx_capi.cu -- the C ABI declared in include/gcx_norm.h: context, residency, kernel drivers.
// Host side is plain C++/CUDA runtime; no torch types cross this boundary.  No CPU fallback.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <stdarg.h>
#include <stdlib.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <unistd.h>
#include <fcntl.h>
#include <errno.h>
#include <sys/stat.h>
#include <sys/mman.h>

#include <algorithm>
#include <functional>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/gcx_norm.h"
#include "gcx_kernels.cuh"
#include "gcx_i8gemm.cuh"

using namespace gcx;

// ---------------------------------------------------------------------------------------------------
// errors
// ---------------------------------------------------------------------------------------------------
static thread_local std::string g_err;
static int fail(const char* fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_err = buf;
    return 1;
}
#define CK(call)                                                                                        \
    do {                                                                                                \
        cudaError_t e__ = (call);                                                                       \
        if (e__ != cudaSuccess) return fail("%s failed at %s:%d: %s", #call, __FILE__, __LINE__, cudaGetErrorString(e__)); \
    } while (0)
#define CKR(call)                  \
    do {                           \
        int r__ = (call);          \
        if (r__ != 0) return r__;  \
    } while (0)

// ---------------------------------------------------------------------------------------------------
// NCCL through dlopen (torch bundles libnccl.so.2; the library itself has no link-time dependency)
// ---------------------------------------------------------------------------------------------------
typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
enum { ncclInt32 = 2, ncclUint32 = 3, ncclUint64 = 5, ncclFloat32 = 7, ncclFloat64 = 8, ncclUint8 = 1 };
enum { ncclSum = 0, ncclProd = 1, ncclMax = 2, ncclMin = 3 };
struct NcclApi {
    void* h = nullptr;
    bool ready = false;
    int (*GetUniqueId)(ncclUniqueId*) = nullptr;
    int (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    int (*CommDestroy)(ncclComm_t) = nullptr;
    int (*AllGather)(const void*, void*, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
};
static NcclApi g_nccl;
static std::mutex g_loader_mutex;          // contexts may be created from several host threads
static int nccl_load() {
    std::lock_guard<std::mutex> lock(g_loader_mutex);
    if (g_nccl.ready) return 0;
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* nm : names) {
        g_nccl.h = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
        if (g_nccl.h) break;
    }
    if (!g_nccl.h) return fail("cannot dlopen libnccl.so.2 (import torch first, or set LD_LIBRARY_PATH): %s", dlerror());
#define SYM(f)                                                          \
    *(void**)(&g_nccl.f) = dlsym(g_nccl.h, "nccl" #f);                   \
    if (!g_nccl.f) return fail("libnccl: missing symbol nccl" #f);
    SYM(GetUniqueId) SYM(CommInitRank) SYM(CommDestroy) SYM(AllGather) SYM(AllReduce) SYM(GetErrorString)
#undef SYM
    g_nccl.ready = true;
    return 0;
}
#define NK(call)                                                                                           \
    do {                                                                                                   \
        int e__ = (call);                                                                                  \
        if (e__ != 0) return fail("%s failed at %s:%d: %s", #call, __FILE__, __LINE__, g_nccl.GetErrorString(e__)); \
    } while (0)

// ---------------------------------------------------------------------------------------------------
// cuBLAS through dlopen (row N4 only: three plain float64 GEMMs per correlation matrix)
// ---------------------------------------------------------------------------------------------------
typedef struct cublasContext* cublasHandle_t;
enum { CUBLAS_OP_N_ = 0, CUBLAS_OP_T_ = 1 };
struct CublasApi {
    void* h = nullptr;
    bool ready = false;
    int (*Create)(cublasHandle_t*) = nullptr;
    int (*Destroy)(cublasHandle_t) = nullptr;
    int (*SetStream)(cublasHandle_t, cudaStream_t) = nullptr;
    int (*Dgemm)(cublasHandle_t, int, int, int, int, int, const double*, const double*, int, const double*, int, const double*, double*, int) = nullptr;
    int (*Dsyrk)(cublasHandle_t, int, int, int, int, const double*, const double*, int, const double*, double*, int) = nullptr;
    int (*GemmEx)(cublasHandle_t, int, int, int, int, int, const void*, const void*, int, int, const void*, int, int, const void*, void*, int, int,
                  int, int) = nullptr;
};
static CublasApi g_cublas;
static int cublas_load() {
    std::lock_guard<std::mutex> lock(g_loader_mutex);
    if (g_cublas.ready) return 0;
    const char* names[] = {"libcublas.so.12", "libcublas.so"};
    for (const char* nm : names) {
        g_cublas.h = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
        if (g_cublas.h) break;
    }
    if (!g_cublas.h) return fail("cannot dlopen libcublas.so.12: %s", dlerror());
    *(void**)(&g_cublas.Create) = dlsym(g_cublas.h, "cublasCreate_v2");
    *(void**)(&g_cublas.Destroy) = dlsym(g_cublas.h, "cublasDestroy_v2");
    *(void**)(&g_cublas.SetStream) = dlsym(g_cublas.h, "cublasSetStream_v2");
    *(void**)(&g_cublas.Dgemm) = dlsym(g_cublas.h, "cublasDgemm_v2");
    *(void**)(&g_cublas.Dsyrk) = dlsym(g_cublas.h, "cublasDsyrk_v2");
    *(void**)(&g_cublas.GemmEx) = dlsym(g_cublas.h, "cublasGemmEx");
    if (!g_cublas.Create || !g_cublas.Destroy || !g_cublas.SetStream || !g_cublas.Dgemm || !g_cublas.Dsyrk || !g_cublas.GemmEx)
        return fail("libcublas: missing symbol");
    g_cublas.ready = true;
    return 0;
}
#define BK(call)                                                                                  \
    do {                                                                                          \
        int e__ = (call);                                                                         \
        if (e__ != 0) return fail("%s failed at %s:%d: cuBLAS status %d", #call, __FILE__, __LINE__, e__); \
    } while (0)

// ---------------------------------------------------------------------------------------------------
// context
// ---------------------------------------------------------------------------------------------------
struct EvPair {
    int cls;
    cudaEvent_t a, b;
};

struct gcx_ctx {
    int device = 0, sm_count = 0;
    cudaStream_t stream = nullptr;
    cudaStream_t copy_stream = nullptr;          // async D2H of results, overlapped with compute
    cudaEvent_t ev_ready = nullptr, ev_copied = nullptr;
    bool copy_pending = false;
    bool last_sym = false;       // the last ic/kr run used the upper-triangle passes
    // map
    int64_t n = 0, ld = 0, row0 = 0, nloc = 0, nv = 0;
    float* A = nullptr;
    float* Out = nullptr;
    float* Out2 = nullptr;       // second result buffer of the fused apply + VC pass
    uint8_t* keep2 = nullptr;    // its VC mask
    MinMax* mm2 = nullptr;
    // vectors (float64, nv each) carved from one slab
    double* slab = nullptr;
    double *vz = nullptr, *vq = nullptr, *vB = nullptr, *vscale = nullptr, *vrowsum = nullptr, *vavg = nullptr;
    double2* vavr = nullptr;             // (vavg, 1 / vavg) per distance, rebuilt before every distance-frequency apply
    KrVecs kr{};
    float* csum32 = nullptr;
    int32_t* zero_count = nullptr;
    int32_t* neg_flag = nullptr;           // 1: the row has negative entries (its float32 emptiness test is re-checked on the host)
    uint8_t* keep = nullptr;
    bool have_mask = false, have_rowsum = false, have_scale = false;
    // matvec workspace
    double* ws = nullptr;
    unsigned int* ticket = nullptr;
    // symmetric (upper-triangle) passes: single rank, exactly symmetric map
    int sym_enable = 1;          // GCX_SYM=0 disables
    int sym_state = -1;          // -1 unknown, 0 not symmetric, 1 symmetric
    double* rowpart = nullptr;   // [nch][ld]: one partial per (panel, row)
    double* colpart = nullptr;   // [grid][nslots][SYM_NG][1024]
    long long* sym_tab_ll = nullptr;      // panel_u0 [nch+1] | cta_u0 [G+1]
    int* sym_tab_i = nullptr;             // cta_kf [G] | cta_kl [G] | panel_cta [nch] | panel_cta_hi [nch]
    int sym_nch = 0;
    int64_t sym_bytes = 0;
    long long* sym_dbg = nullptr;          // per-CTA main-loop cycles (GCX_SYM_DEBUG=1)
    long long* sym_stamps = nullptr;       // [G][16] phase timestamps of one launch of k_ic_pass_symd (GCX_SYM_DEBUG=1)
    int sym_dbg_launch = 20;               // GCX_SYM_DEBUG_LAUNCH
    double* colpart_dyn = nullptr;         // [ndyn][SYM_NG][1024] column partials of the dynamic chunks
    unsigned int* sym_dyn_counter = nullptr;
    int sym_ndyn = 0;
    int sym_blocks = 1;                   // sharded: circulant assignment of off-diagonal blocks (needs the aligned row partition)
    int sym_kfold_lo = -1;
    int sym_balance = 1;                 // move the cuts of the block pairs until the ranks' shares of a pass are level
    int sym_nfoldk = 0;
    int assume_symmetric = 0;             // sharded runs cannot test symmetry locally: the caller vouches for it
    int exchange_allreduce = 0;           // custom (unequal) row partition: vectors are exchanged by zero-fill + allreduce
    int sym_slots = 0, sym_grid = 0;
    void* scratch = nullptr;     // grow-only arena for the per-diagonal statistics (avoids cudaMalloc/cudaFree per call)
    size_t scratch_bytes = 0;
    double* part = nullptr;      // per-CTA partials of the fused IC pass (3 x grid)
    int ic_fused = 1;            // 1: one cooperative launch per IC pass (k_ic_pass); 0: k_matvec + k_ic_epilogue
    int ic_persist = 1;          // deferred-scalar pass: the whole solve in ONE cooperative launch (GCX_IC_PERSIST=0: one launch per pass)
    int sym_l2_hint = 1;         // evict-first L2 policy on the map stream of the upper-triangle pass (GCX_SYM_L2HINT)
    int ic_defer = 1;            // upper-triangle path: deferred-scalar pass k_ic_pass_symd (GCX_IC_DEFER=0: literal-order k_ic_pass_sym_tma)
    double* ic_ring = nullptr;   // [3][nv] t' ring + [2][nv] u ring of the deferred-scalar pass
    int last_matvecs_issued = 0; // products the last ic/kr run really launched (the deferred pass runs one more than passes + 1)
    // peer exchange of the sharded symmetric pass (NVLink stores into peer-mapped buffers; cudaIpc or same-process P2P)
    int xchg_enable = 1;         // option "peer_exchange" / GCX_PEER_EXCHANGE=0: keep the NCCL allreduce
    bool xchg_ok = false;
    void* xchg_buf = nullptr;    // [world][G] flags | [2][world][nv] partial vectors
    void* xchg_peer[8] = {nullptr};
    bool xchg_ipc[8] = {false};
    unsigned int xchg_seq = 1;   // flag value of the next run's launch 0 (same on every rank: SPMD call sequence)
    int ic_grid = 0, ic_grid_tma = 0;
    int corr_own_gemm = 1;  // int8 slice / indicator products on this library's tcgen05 kernel (k_i8gemm_nt); 0: cublasGemmEx (option "corr_own_gemm", GCX_CORR_OWN_GEMM)
    unsigned int i8_cfg_mask = 0;
    bool selfvar_ready = false;
    int corr_int_path = 1;  // variant 1: exact int8 slice products for Sx and Sxy when the map is integer-valued (option "corr_int_path")
    int corr_last_int = 0;
    int corr_variant = 1;   // correlation: 1 = bf16 tensor-core counts + DSYRK + DGEMM, 0 = three DGEMMs (GCX_CORR_VARIANT / option "corr_variant")
    int sd_variant = 1;   // stationary distribution: 1 = TMA-bulk transposed product, 0 = LDG (GCX_SD_VARIANT / option "sd_variant")
    unsigned int mv_cfg_mask = 0;   // TMA matvec variants whose shared-memory attribute has been set on this device
    int mvt_occ = -1;               // CTAs/SM of the transposed TMA product on this device (-1: not configured yet)
    int mv_variant = 7;   // TMA-bulk, R=4 rows per band, 6 stages, 2 CTAs/SM: best of profiles/r01_matvec_variants.txt
    // states
    IcState* ic_state = nullptr;
    KrState* kr_state = nullptr;
    SdState* sd_state = nullptr;
    MinMax* mm = nullptr;
    unsigned long long* csum_dev = nullptr;
    void* h_state = nullptr;   // pinned, 4 KB
    // staging
    void* stage[2] = {nullptr, nullptr};
    cudaEvent_t stage_ev[2] = {nullptr, nullptr};
    size_t stage_bytes = 0;
    // dist
    int rank = 0, world = 1;
    int64_t rpr = 0;   // rows per rank (allgather count)
    ncclComm_t comm = nullptr;
    cublasHandle_t blas = nullptr;          // created on the first correlation run
    // measurement
    bool profile = false;
    std::vector<EvPair> evs;
    std::vector<cudaEvent_t> ev_pool;
    int64_t prof_launches[GCX_NCLASS] = {0};
    double prof_ms[GCX_NCLASS] = {0};
    int64_t launches = 0;
    cudaEvent_t run_a = nullptr, run_b = nullptr;
    cudaEvent_t tm_a = nullptr, tm_b = nullptr;
    double last_run_ms = 0;
};

static const size_t STAGE_BYTES = 64ull << 20;

static int set_dev(gcx_ctx* c) {
    CK(cudaSetDevice(c->device));
    return 0;
}

static int prof_fold(gcx_ctx* c) {
    if (c->evs.empty()) return 0;
    CK(cudaStreamSynchronize(c->stream));
    for (auto& p : c->evs) {
        float ms = 0;
        CK(cudaEventElapsedTime(&ms, p.a, p.b));
        c->prof_ms[p.cls] += ms;
        c->prof_launches[p.cls] += 1;
        c->ev_pool.push_back(p.a);
        c->ev_pool.push_back(p.b);
    }
    c->evs.clear();
    return 0;
}

static int ev_get(gcx_ctx* c, cudaEvent_t* e) {
    if (!c->ev_pool.empty()) {
        *e = c->ev_pool.back();
        c->ev_pool.pop_back();
        return 0;
    }
    CK(cudaEventCreate(e));
    return 0;
}

struct ProfScope {
    gcx_ctx* c;
    int cls;
    bool on;
    cudaEvent_t a = nullptr, b = nullptr;
    ProfScope(gcx_ctx* c_, int cls_) : c(c_), cls(cls_), on(c_->profile) {
        if (on) {
            if (ev_get(c, &a) || ev_get(c, &b)) { on = false; return; }
            cudaEventRecord(a, c->stream);
        }
    }
    ~ProfScope() {
        c->launches += 1;
        if (on) {
            cudaEventRecord(b, c->stream);
            c->evs.push_back({cls, a, b});
        }
    }
};

// Small device->host reads (convergence state, min/max, counters) go through MAPPED pinned memory written by this
// tiny kernel, never through a D2H memcpy: a copy-engine transfer would queue behind any large async download in
// flight on the copy stream (head-of-line blocking) and stall the caller for its whole duration.
__global__ void k_publish_state(const int* __restrict__ src, volatile int* dst, int words) {
    for (int i = threadIdx.x; i < words; i += blockDim.x) dst[i] = src[i];
    __threadfence_system();
}

static int read_small(gcx_ctx* c, const void* dev, size_t bytes, void* host_dst) {
    if (bytes > 1024 || bytes % 4) return fail("read_small: bad size");
    volatile int* slot = (volatile int*)((char*)c->h_state + 3072);
    k_publish_state<<<1, 64, 0, c->stream>>>((const int*)dev, slot, (int)(bytes / 4));
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(c->stream));
    memcpy(host_dst, (const void*)slot, bytes);
    return 0;
}

// ---------------------------------------------------------------------------------------------------
// library / context
// ---------------------------------------------------------------------------------------------------
extern "C" const char* gcx_last_error(void) { return g_err.c_str(); }
extern "C" int gcx_version(void) { return 100; }

extern "C" int gcx_device_count(int* count) {
    CK(cudaGetDeviceCount(count));
    return 0;
}

extern "C" int gcx_ctx_destroy(gcx_ctx* c);
static void stage_give(void* p);

static int ctx_init(gcx_ctx* c, int device) {
    c->device = device;
    CK(cudaSetDevice(device));
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10) return fail("device %d is sm_%d%d; libgcxnorm is built for sm_100a (B200) only", device, prop.major, prop.minor);
    c->sm_count = prop.multiProcessorCount;
    CK(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    CK(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
    CK(cudaEventCreateWithFlags(&c->ev_ready, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&c->ev_copied, cudaEventDisableTiming));
    CK(cudaMalloc(&c->ic_state, sizeof(IcState)));
    CK(cudaMalloc(&c->kr_state, sizeof(KrState)));
    CK(cudaMalloc(&c->sd_state, sizeof(SdState)));
    CK(cudaMalloc(&c->mm, sizeof(MinMax)));
    CK(cudaMalloc(&c->mm2, sizeof(MinMax)));
    CK(cudaMalloc(&c->csum_dev, sizeof(unsigned long long)));
    CK(cudaMalloc(&c->ticket, sizeof(unsigned int)));
    CK(cudaMemset(c->ticket, 0, sizeof(unsigned int)));
    CK(cudaMalloc(&c->ws, sizeof(double) * (size_t)c->sm_count * 8 * 2 * 16));
    CK(cudaMalloc(&c->part, sizeof(double) * 3 * (size_t)c->sm_count * 8));
    {
        int occ = 0, coop = 0;
        CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_ic_pass<4, 4>, MV_THREADS, 0));
        CK(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, device));
        c->ic_grid = c->sm_count * std::min(occ, 4);
        if (!coop || occ < 1) c->ic_fused = 0;
        const char* f = getenv("GCX_IC_FUSED");
        if (f) c->ic_fused = atoi(f) && coop && occ >= 1;     // 0 split, 1 fused LDG, 2 fused TMA-bulk (default)
        if (c->ic_fused) {
            const int smem = 6 * 4 * MV_THREADS * (int)sizeof(float4);
            CK(cudaFuncSetAttribute(k_ic_pass_tma<4, 6, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
            int occ2 = 0;
            CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ2, k_ic_pass_tma<4, 6, 2>, TMA_THREADS, smem));
            c->ic_grid_tma = c->sm_count * std::min(occ2, 2);
            if (!f) c->ic_fused = occ2 >= 1 ? 2 : 1;
            else if (atoi(f) == 2 && occ2 < 1) c->ic_fused = 1;
            else if (atoi(f) == 2) c->ic_fused = 2;
        }
    }
    CK(cudaHostAlloc(&c->h_state, 4096, cudaHostAllocMapped));     // UVA: the same pointer is valid on the device
    CK(cudaEventCreate(&c->run_a));
    CK(cudaEventCreate(&c->run_b));
    CK(cudaEventCreate(&c->tm_a));
    CK(cudaEventCreate(&c->tm_b));
    const char* v = getenv("GCX_MATVEC_VARIANT");
    if (v) c->mv_variant = atoi(v);
    if (const char* sv = getenv("GCX_SD_VARIANT")) c->sd_variant = atoi(sv);
    if (const char* cv = getenv("GCX_CORR_VARIANT")) c->corr_variant = atoi(cv);
    if (const char* cv = getenv("GCX_CORR_OWN_GEMM")) c->corr_own_gemm = atoi(cv);
    const char* sy = getenv("GCX_SYM");
    if (sy) c->sym_enable = atoi(sy);
    if (c->ic_fused != 2) c->sym_enable = c->sym_enable && (c->ic_fused != 1);   // the fused LDG pass has no symmetric twin
    {
        const int smem = SYM_STAGES * SYM_R * MV_THREADS * (int)sizeof(float4);
        CK(cudaFuncSetAttribute(k_matvec_sym_tma<SYM_STAGES, SYM_OCC>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        CK(cudaFuncSetAttribute(k_ic_pass_sym_tma<SYM_STAGES, SYM_OCC, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        CK(cudaFuncSetAttribute(k_ic_pass_sym_tma<SYM_STAGES, SYM_OCC, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        CK(cudaFuncSetAttribute(k_ic_pass_symd<SYM_STAGES, SYM_OCC, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        CK(cudaFuncSetAttribute(k_ic_pass_symd<SYM_STAGES, SYM_OCC, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        if (const char* e = getenv("GCX_IC_DEFER")) c->ic_defer = atoi(e);
        if (const char* e = getenv("GCX_IC_PERSIST")) c->ic_persist = atoi(e);
        if (const char* e = getenv("GCX_SYM_L2HINT")) c->sym_l2_hint = atoi(e);
        if (const char* e = getenv("GCX_PEER_EXCHANGE")) c->xchg_enable = atoi(e);
        int occ3 = 0, occ4 = 0;
        CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ3, k_ic_pass_sym_tma<SYM_STAGES, SYM_OCC, true>, SYM_THREADS, smem));
        CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ4, k_ic_pass_symd<SYM_STAGES, SYM_OCC, true>, SYM_THREADS, smem));
        if (occ4 < occ3) c->ic_defer = 0;          // both passes share one grid (tables, partial buffers)
        c->sym_grid = c->sm_count * std::min(occ3, SYM_OCC);
        if (occ3 < 1) c->sym_enable = 0;
    }
    return 0;
}

extern "C" int gcx_ctx_create(int device, gcx_ctx** out) {
    if (!out) return fail("gcx_ctx_create: null output pointer");
    *out = nullptr;
    int cnt = 0;
    CK(cudaGetDeviceCount(&cnt));
    if (cnt <= 0) return fail("no CUDA device visible: libgcxnorm has no CPU fallback");
    if (device < 0 || device >= cnt) return fail("device %d out of range (%d visible)", device, cnt);
    gcx_ctx* c = new gcx_ctx();
    if (ctx_init(c, device)) {
        gcx_ctx_destroy(c);       // releases whatever was created; the error text stays
        return 1;
    }
    *out = c;
    return 0;
}

// tables and partial buffers of the upper-triangle pass (rebuilt by sym_ready on the next run)
static void free_sym(gcx_ctx* c) {
    cudaFree(c->rowpart); c->rowpart = nullptr;
    cudaFree(c->colpart); c->colpart = nullptr;
    cudaFree(c->sym_tab_ll); c->sym_tab_ll = nullptr;
    cudaFree(c->sym_tab_i); c->sym_tab_i = nullptr;
    cudaFree(c->sym_dbg); c->sym_dbg = nullptr;
    cudaFree(c->sym_stamps); c->sym_stamps = nullptr;
    cudaFree(c->colpart_dyn); c->colpart_dyn = nullptr;
    cudaFree(c->sym_dyn_counter); c->sym_dyn_counter = nullptr;
    c->sym_ndyn = 0;
}

static void xchg_free(gcx_ctx* c) {
    for (int q = 0; q < 8; ++q) {
        if (c->xchg_peer[q] && c->xchg_ipc[q]) cudaIpcCloseMemHandle(c->xchg_peer[q]);
        c->xchg_peer[q] = nullptr;
        c->xchg_ipc[q] = false;
    }
    cudaFree(c->xchg_buf); c->xchg_buf = nullptr;
    c->xchg_ok = false;
}

static void free_map(gcx_ctx* c) {
    xchg_free(c);
    cudaFree(c->ic_ring); c->ic_ring = nullptr;
    cudaFree(c->A); c->A = nullptr;
    cudaFree(c->Out); c->Out = nullptr;
    cudaFree(c->Out2); c->Out2 = nullptr;
    cudaFree(c->keep2); c->keep2 = nullptr;
    free_sym(c);
    c->sym_state = -1;
    cudaFree(c->slab); c->slab = nullptr;
    cudaFree(c->csum32); c->csum32 = nullptr;
    cudaFree(c->zero_count); c->zero_count = nullptr;
    cudaFree(c->neg_flag); c->neg_flag = nullptr;
    cudaFree(c->keep); c->keep = nullptr;
    c->have_mask = c->have_rowsum = c->have_scale = false;
}

extern "C" int gcx_ctx_destroy(gcx_ctx* c) {
    if (!c) return 0;
    cudaSetDevice(c->device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    if (c->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(c->comm);
    if (c->blas && g_cublas.Destroy) g_cublas.Destroy(c->blas);
    free_map(c);
    for (auto& p : c->evs) { cudaEventDestroy(p.a); cudaEventDestroy(p.b); }
    for (auto e : c->ev_pool) cudaEventDestroy(e);
    cudaFree(c->ic_state); cudaFree(c->kr_state); cudaFree(c->sd_state); cudaFree(c->mm); cudaFree(c->mm2); cudaFree(c->csum_dev); cudaFree(c->ticket); cudaFree(c->ws); cudaFree(c->part); cudaFree(c->scratch);
    if (c->h_state) cudaFreeHost(c->h_state);
    for (int i = 0; i < 2; ++i) {
        if (c->stage[i]) stage_give(c->stage[i]);
        if (c->stage_ev[i]) cudaEventDestroy(c->stage_ev[i]);
    }
    for (cudaEvent_t e : {c->run_a, c->run_b, c->tm_a, c->tm_b})
        if (e) cudaEventDestroy(e);
    if (c->copy_stream) cudaStreamSynchronize(c->copy_stream);
    for (cudaEvent_t e : {c->ev_ready, c->ev_copied})
        if (e) cudaEventDestroy(e);
    if (c->copy_stream) cudaStreamDestroy(c->copy_stream);
    if (c->stream) cudaStreamDestroy(c->stream);
    cudaGetLastError();           // a half-built context may have produced sticky-free errors above
    delete c;
    return 0;
}

extern "C" int gcx_sync(gcx_ctx* c) {
    CKR(set_dev(c));
    CK(cudaStreamSynchronize(c->stream));
    return 0;
}

extern "C" int gcx_device_info(gcx_ctx* c, int* sm_count, int64_t* hbm_total, int64_t* hbm_free) {
    CKR(set_dev(c));
    size_t f = 0, t = 0;
    CK(cudaMemGetInfo(&f, &t));
    if (sm_count) *sm_count = c->sm_count;
    if (hbm_total) *hbm_total = (int64_t)t;
    if (hbm_free) *hbm_free = (int64_t)f;
    return 0;
}

extern "C" int gcx_host_alloc(int64_t bytes, void** ptr) {
    CK(cudaHostAlloc(ptr, (size_t)bytes, cudaHostAllocDefault));
    return 0;
}
extern "C" int gcx_host_free(void* ptr) {
    CK(cudaFreeHost(ptr));
    return 0;
}

// ---------------------------------------------------------------------------------------------------
// map residency
// ---------------------------------------------------------------------------------------------------
static int64_t round_up(int64_t x, int64_t m) { return (x + m - 1) / m * m; }

static int map_alloc(gcx_ctx* c, int64_t n, int64_t row0, int64_t nloc);
static int xchg_setup(gcx_ctx* c);
static int allgather_f64(gcx_ctx* c, double* vec);
static int allgather_i32(gcx_ctx* c, int32_t* vec);

extern "C" int gcx_map_create(gcx_ctx* c, int64_t n, int64_t row0, int64_t nloc) {
    if (!c) return fail("gcx_map_create: null context");
    const int rc = map_alloc(c, n, row0, nloc);
    if (rc) {
        free_map(c);              // never leave a half-allocated map behind (need_map keys on c->A)
        cudaGetLastError();       // e.g. an out-of-memory error must not surface again at the next launch check
    }
    return rc;
}

static int map_alloc(gcx_ctx* c, int64_t n, int64_t row0, int64_t nloc) {
    CKR(set_dev(c));
    if (n <= 0 || nloc <= 0 || row0 < 0 || row0 + nloc > n) return fail("bad map geometry n=%lld row0=%lld nloc=%lld", (long long)n, (long long)row0, (long long)nloc);
    if (n > 0x7ffffff0LL) return fail("n too large");
    CK(cudaStreamSynchronize(c->stream));
    free_map(c);
    c->n = n; c->row0 = row0; c->nloc = nloc;
    c->ld = round_up(n, 32);
    c->rpr = (n + c->world - 1) / c->world;
    if (c->world > 1 && !c->exchange_allreduce) {
        // the allgather exchange sends slot `rank` of ceil(n/world) entries: any other row block would be gathered into the wrong place
        const int64_t want0 = std::min<int64_t>(n, (int64_t)c->rank * c->rpr), wantn = std::max<int64_t>(0, std::min<int64_t>(c->rpr, n - want0));
        if (row0 != want0 || nloc != wantn)
            return fail("row block [%lld, +%lld) of rank %d is not the equal ceil(n/world) block [%lld, +%lld) the allgather exchange assumes: "
                        "set the option \"exchange_allreduce\" on every rank for custom partitions (gcx_partition_rows, equal-area cuts)",
                        (long long)row0, (long long)nloc, c->rank, (long long)want0, (long long)wantn);
    }
    c->nv = round_up(std::max<int64_t>(c->ld, c->rpr * c->world), 32) + 32;
    CK(cudaMalloc(&c->A, sizeof(float) * (size_t)nloc * c->ld));
    CK(cudaMemsetAsync(c->A, 0, sizeof(float) * (size_t)nloc * c->ld, c->stream));
    const int NVEC = 15;
    CK(cudaMalloc(&c->slab, sizeof(double) * (size_t)c->nv * NVEC));
    CK(cudaMemsetAsync(c->slab, 0, sizeof(double) * (size_t)c->nv * NVEC, c->stream));
    double* p = c->slab;
    auto take = [&]() { double* r = p; p += c->nv; return r; };
    c->vz = take(); c->vq = take(); c->vB = take(); c->vscale = take(); c->vrowsum = take(); c->vavg = take();
    c->kr.x = take(); c->kr.v = take(); c->kr.rk = take(); c->kr.y = take(); c->kr.Z = take(); c->kr.p = take(); c->kr.w = take();
    c->kr.z = c->vz;
    c->vavr = reinterpret_cast<double2*>(take());      // two slots
    take();
    CK(cudaMalloc(&c->csum32, sizeof(float) * (size_t)c->nv));
    CK(cudaMemsetAsync(c->csum32, 0, sizeof(float) * (size_t)c->nv, c->stream));
    CK(cudaMalloc(&c->zero_count, sizeof(int32_t) * (size_t)c->nv));
    CK(cudaMemsetAsync(c->zero_count, 0, sizeof(int32_t) * (size_t)c->nv, c->stream));
    CK(cudaMalloc(&c->neg_flag, sizeof(int32_t) * (size_t)c->nv));
    CK(cudaMemsetAsync(c->neg_flag, 0, sizeof(int32_t) * (size_t)c->nv, c->stream));
    CK(cudaMalloc(&c->keep, (size_t)c->nv));
    CK(cudaMemsetAsync(c->keep, 0, (size_t)c->nv, c->stream));
    CK(cudaMalloc(&c->ic_ring, sizeof(double) * (size_t)c->nv * 5));
    CK(cudaMemsetAsync(c->ic_ring, 0, sizeof(double) * (size_t)c->nv * 5, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    if (c->world > 1 && c->xchg_enable) CKR(xchg_setup(c));
    return 0;
}

static int need_map(gcx_ctx* c) {
    if (!c || !c->A) return fail("no resident map: call gcx_map_create first");
    return set_dev(c);
}

static int fence_copy(gcx_ctx* c);

static int ensure_out(gcx_ctx* c) {
    if (!c->Out) CK(cudaMalloc(&c->Out, sizeof(float) * (size_t)c->nloc * c->ld));
    return 0;
}

static int ensure_scratch(gcx_ctx* c, size_t bytes) {
    if (bytes <= c->scratch_bytes) return 0;
    CK(cudaStreamSynchronize(c->stream));
    if (c->scratch) CK(cudaFree(c->scratch));
    c->scratch = nullptr;
    c->scratch_bytes = 0;
    CK(cudaMalloc(&c->scratch, bytes));
    c->scratch_bytes = bytes;
    return 0;
}

// Pinned staging buffers are expensive to create (cudaHostAlloc of 64 MB: tens of milliseconds) and every drop-in call makes a
// fresh context, so released stages are parked in a small process-wide pool (cudaHostAllocPortable: usable from any device).
static std::mutex g_stage_mutex;
static std::vector<void*> g_stage_pool;
static int stage_take(void** p) {
    {
        std::lock_guard<std::mutex> lock(g_stage_mutex);
        if (!g_stage_pool.empty()) { *p = g_stage_pool.back(); g_stage_pool.pop_back(); return 0; }
    }
    CK(cudaHostAlloc(p, STAGE_BYTES, cudaHostAllocPortable));
    return 0;
}
static void stage_give(void* p) {
    std::lock_guard<std::mutex> lock(g_stage_mutex);
    if (g_stage_pool.size() < 16) g_stage_pool.push_back(p);
    else cudaFreeHost(p);
}

static int ensure_stage(gcx_ctx* c) {
    if (c->stage[0]) return 0;
    c->stage_bytes = STAGE_BYTES;
    for (int i = 0; i < 2; ++i) {
        CKR(stage_take(&c->stage[i]));
        CK(cudaEventCreateWithFlags(&c->stage_ev[i], cudaEventDisableTiming));
    }
    return 0;
}

static bool is_pinned(const void* p) {
    cudaPointerAttributes a;
    cudaError_t e = cudaPointerGetAttributes(&a, p);
    if (e != cudaSuccess) { cudaGetLastError(); return false; }
    return a.type == cudaMemoryTypeHost || a.type == cudaMemoryTypeManaged;
}

static void par_copy_rows(char* dst, size_t dpitch, const char* src, size_t spitch, size_t width, int64_t rows) {
    unsigned hw = std::thread::hardware_concurrency();
    static const unsigned cap = [] { const char* e = getenv("GCX_COPY_THREADS"); return e ? (unsigned)std::max(1, atoi(e)) : 16u; }();
    int T = (int)std::min<int64_t>(std::max(1u, std::min(hw, cap)), rows);
    if ((size_t)rows * width < (4u << 20)) T = 1;
    auto work = [=](int t) {
        const int64_t a = rows * t / T, b = rows * (t + 1) / T;
        for (int64_t r = a; r < b; ++r) memcpy(dst + r * dpitch, src + r * spitch, width);
    };
    if (T == 1) { work(0); return; }
    std::vector<std::thread> th;
    for (int t = 1; t < T; ++t) th.emplace_back(work, t);
    work(0);
    for (auto& x : th) x.join();
}

extern "C" int gcx_map_upload(gcx_ctx* c, const float* host, int64_t host_ld) {
    CKR(need_map(c));
    CKR(fence_copy(c));
    c->sym_state = -1;
    c->have_rowsum = false;      // everything derived from the previous contents is stale
    c->have_scale = false;
    if (host_ld < c->n) return fail("host_ld < n");
    const size_t width = sizeof(float) * (size_t)c->n;
    if (is_pinned(host)) {
        if (host_ld == c->ld)
            CK(cudaMemcpyAsync(c->A, host, sizeof(float) * (size_t)c->nloc * c->ld, cudaMemcpyHostToDevice, c->stream));
        else
            CK(cudaMemcpy2DAsync(c->A, sizeof(float) * c->ld, host, sizeof(float) * host_ld, width, (size_t)c->nloc, cudaMemcpyHostToDevice, c->stream));
        CK(cudaStreamSynchronize(c->stream));
        return 0;
    }
    CKR(ensure_stage(c));
    const int64_t rows_per = std::max<int64_t>(1, (int64_t)(c->stage_bytes / width));
    int64_t r0 = 0;
    int it = 0;
    while (r0 < c->nloc) {
        const int b = it & 1;
        const int64_t rows = std::min(rows_per, c->nloc - r0);
        if (it >= 2) CK(cudaEventSynchronize(c->stage_ev[b]));
        par_copy_rows((char*)c->stage[b], width, (const char*)(host + r0 * host_ld), sizeof(float) * host_ld, width, rows);
        CK(cudaMemcpy2DAsync(c->A + r0 * c->ld, sizeof(float) * c->ld, c->stage[b], width, width, (size_t)rows, cudaMemcpyHostToDevice, c->stream));
        CK(cudaEventRecord(c->stage_ev[b], c->stream));
        r0 += rows;
        ++it;
    }
    CK(cudaStreamSynchronize(c->stream));
    return 0;
}

extern "C" int gcx_map_download(gcx_ctx* c, int which, float* host, int64_t host_ld) {
    CKR(need_map(c));
    const float* src = which == 0 ? c->A : (which == 2 ? c->Out2 : c->Out);
    if (!src) return fail("no output map resident");
    if (host_ld < c->n) return fail("host_ld < n");
    const size_t width = sizeof(float) * (size_t)c->n;
    if (is_pinned(host)) {
        if (host_ld == c->ld)
            CK(cudaMemcpyAsync(host, src, sizeof(float) * (size_t)c->nloc * c->ld, cudaMemcpyDeviceToHost, c->stream));
        else
            CK(cudaMemcpy2DAsync(host, sizeof(float) * host_ld, src, sizeof(float) * c->ld, width, (size_t)c->nloc, cudaMemcpyDeviceToHost, c->stream));
        CK(cudaStreamSynchronize(c->stream));
        return 0;
    }
    CKR(ensure_stage(c));
    const int64_t rows_per = std::max<int64_t>(1, (int64_t)(c->stage_bytes / width));
    int64_t r0 = 0, prev_r0 = -1, prev_rows = 0;
    int it = 0;
    while (r0 < c->nloc || prev_r0 >= 0) {
        const int b = it & 1;
        int64_t rows = 0;
        if (r0 < c->nloc) {
            rows = std::min(rows_per, c->nloc - r0);
            CK(cudaMemcpy2DAsync(c->stage[b], width, src + r0 * c->ld, sizeof(float) * c->ld, width, (size_t)rows, cudaMemcpyDeviceToHost, c->stream));
            CK(cudaEventRecord(c->stage_ev[b], c->stream));
        }
        if (prev_r0 >= 0) {
            CK(cudaEventSynchronize(c->stage_ev[b ^ 1]));
            par_copy_rows((char*)(host + prev_r0 * host_ld), sizeof(float) * host_ld, (const char*)c->stage[b ^ 1], width, width, prev_rows);
        }
        if (rows > 0) { prev_r0 = r0; prev_rows = rows; r0 += rows; } else { prev_r0 = -1; }
        ++it;
    }
    return 0;
}

// ---------------------------------------------------------------------------------------------------
// streaming residency: rows arrive in chunks -- from a .npbin file read by a pool of pread threads, from a
// gzip stream, from an HDF5 dataset -- and every chunk's row statistics are computed as it lands, so the mask pass overlaps the
// upload and the map is read from HBM one time less.
// ---------------------------------------------------------------------------------------------------
static void par_for(int64_t items, size_t bytes_total, const std::function<void(int64_t, int64_t)>& work) {
    unsigned hw = std::thread::hardware_concurrency();
    static const unsigned cap = [] { const char* e = getenv("GCX_COPY_THREADS"); return e ? (unsigned)std::max(1, atoi(e)) : 16u; }();
    int T = (int)std::min<int64_t>(std::max(1u, std::min(hw, cap)), items);
    if (bytes_total < (4u << 20)) T = 1;
    if (T <= 1) { work(0, items); return; }
    std::vector<std::thread> th;
    for (int t = 1; t < T; ++t) th.emplace_back([=, &work] { work(items * t / T, items * (t + 1) / T); });
    work(0, items / T);
    for (auto& x : th) x.join();
}

// row statistics of the local rows [r0, r0 + rows) right behind their host-to-device copy on the stream
static int stats_for_rows(gcx_ctx* c, int64_t r0, int64_t rows) {
    ProfScope ps(c, 3);
    const int grid = (int)std::min<int64_t>(rows, (int64_t)c->sm_count * 8);
    k_row_stats<<<grid, MV_THREADS, 0, c->stream>>>(c->A + r0 * c->ld, c->ld, (int)rows, (int)c->n, c->vrowsum + c->row0 + r0, c->zero_count + c->row0 + r0,
                                                      c->neg_flag + c->row0 + r0);
    CK(cudaGetLastError());
    return 0;
}

extern "C" int gcx_map_upload_rows(gcx_ctx* c, const float* host, int64_t host_ld, int64_t row_begin, int64_t nrows, int with_stats) {
    CKR(need_map(c));
    if (row_begin < 0 || nrows < 0 || row_begin + nrows > c->nloc) return fail("gcx_map_upload_rows: rows [%lld, +%lld) outside the local block of %lld rows", (long long)row_begin, (long long)nrows, (long long)c->nloc);
    if (host_ld < c->n) return fail("host_ld < n");
    if (row_begin == 0) {
        CKR(fence_copy(c));
        c->sym_state = -1; c->have_rowsum = false; c->have_scale = false;
    }
    if (nrows == 0) return 0;
    const size_t width = sizeof(float) * (size_t)c->n;
    if (is_pinned(host)) {
        CK(cudaMemcpy2DAsync(c->A + row_begin * c->ld, sizeof(float) * c->ld, host, sizeof(float) * host_ld, width, (size_t)nrows, cudaMemcpyHostToDevice, c->stream));
        if (with_stats) CKR(stats_for_rows(c, row_begin, nrows));
        CK(cudaStreamSynchronize(c->stream));
        return 0;
    }
    CKR(ensure_stage(c));
    const int64_t rows_per = std::max<int64_t>(1, (int64_t)(c->stage_bytes / width));
    int64_t r0 = 0;
    int it = 0;
    while (r0 < nrows) {
        const int b = it & 1;
        const int64_t rows = std::min(rows_per, nrows - r0);
        if (it >= 2) CK(cudaEventSynchronize(c->stage_ev[b]));
        par_copy_rows((char*)c->stage[b], width, (const char*)(host + r0 * host_ld), sizeof(float) * host_ld, width, rows);
        CK(cudaMemcpy2DAsync(c->A + (row_begin + r0) * c->ld, sizeof(float) * c->ld, c->stage[b], width, width, (size_t)rows, cudaMemcpyHostToDevice, c->stream));
        CK(cudaEventRecord(c->stage_ev[b], c->stream));
        if (with_stats) CKR(stats_for_rows(c, row_begin + r0, rows));
        r0 += rows;
        ++it;
    }
    CK(cudaStreamSynchronize(c->stream));
    return 0;
}

// every row of the local block has been uploaded with with_stats != 0: publish the statistics (allgather when sharded)
extern "C" int gcx_map_upload_finish(gcx_ctx* c) {
    CKR(need_map(c));
    CKR(allgather_f64(c, c->vrowsum));
    CKR(allgather_i32(c, c->zero_count));
    CKR(allgather_i32(c, c->neg_flag));
    c->have_rowsum = true;
    return 0;
}

// The local row block straight from a raw float32 row-major file (.npbin, lib/ccmap.py:233: n x n float32, no header): a pool of
// threads preads row ranges into the two pinned stages while the previous stage travels to the device.  No memmap, no page
// faults on the critical path, no intermediate copy.
extern "C" int gcx_map_upload_file(gcx_ctx* c, const char* path, int64_t file_offset, int with_stats) {
    CKR(need_map(c));
    CKR(fence_copy(c));
    c->sym_state = -1; c->have_rowsum = false; c->have_scale = false;
    const int fd = open(path, O_RDONLY);
    if (fd < 0) return fail("gcx_map_upload_file: cannot open %s: %s", path, strerror(errno));
    struct stat sb;
    const size_t width = sizeof(float) * (size_t)c->n;
    const int64_t base = file_offset + (int64_t)c->row0 * (int64_t)width;
    if (fstat(fd, &sb) != 0 || (int64_t)sb.st_size < base + (int64_t)c->nloc * (int64_t)width) {
        close(fd);
        return fail("gcx_map_upload_file: %s is shorter than rows [%lld, +%lld) of a %lld x %lld float32 map", path, (long long)c->row0, (long long)c->nloc, (long long)c->n, (long long)c->n);
    }
    int rc = ensure_stage(c);
    const int64_t rows_per = std::max<int64_t>(1, (int64_t)(c->stage_bytes / width));
    int64_t r0 = 0;
    int it = 0;
    std::string err;
    std::mutex err_mu;
    while (rc == 0 && r0 < c->nloc) {
        const int b = it & 1;
        const int64_t rows = std::min(rows_per, c->nloc - r0);
        if (it >= 2 && cudaEventSynchronize(c->stage_ev[b]) != cudaSuccess) { rc = fail("cudaEventSynchronize failed"); break; }
        char* dst = (char*)c->stage[b];
        const int64_t off0 = base + r0 * (int64_t)width, total = rows * (int64_t)width;
        // contiguous in the file and in the stage: split by bytes (4 MB granules), not by rows
        const int64_t gran = 4ll << 20, items = (total + gran - 1) / gran;
        par_for(items, (size_t)total, [&](int64_t a, int64_t e) {
            int64_t p = a * gran;
            const int64_t pe = std::min(total, e * gran);
            while (p < pe) {
                const ssize_t got = pread(fd, dst + p, (size_t)(pe - p), (off_t)(off0 + p));
                if (got <= 0) {
                    if (got < 0 && errno == EINTR) continue;
                    std::lock_guard<std::mutex> g(err_mu);
                    err = got == 0 ? "unexpected end of file" : strerror(errno);
                    return;
                }
                p += got;
            }
        });
        if (!err.empty()) { rc = fail("gcx_map_upload_file: reading %s: %s", path, err.c_str()); break; }
        if (cudaMemcpy2DAsync(c->A + r0 * c->ld, sizeof(float) * c->ld, c->stage[b], width, width, (size_t)rows, cudaMemcpyHostToDevice, c->stream) != cudaSuccess ||
            cudaEventRecord(c->stage_ev[b], c->stream) != cudaSuccess) { rc = fail("gcx_map_upload_file: host-to-device copy failed: %s", cudaGetErrorString(cudaGetLastError())); break; }
        if (with_stats) rc = stats_for_rows(c, r0, rows);
        r0 += rows;
        ++it;
    }
    close(fd);
    if (rc) { cudaStreamSynchronize(c->stream); return rc; }
    if (with_stats) CKR(gcx_map_upload_finish(c));
    CK(cudaStreamSynchronize(c->stream));
    return 0;
}

// A result map straight into a raw float32 file (created / sized as needed): device-to-host per stage, and behind it a pool of
// threads copies the stage into a shared mapping of the file.  (pwrite would serialise on the inode lock -- 0.69 s for a new 2.5 GB
// file on tmpfs with 1 or 16 threads -- while page faults on a mapping are taken in parallel: 0.46 s, tools/io_probe.py.)
extern "C" int gcx_map_download_file(gcx_ctx* c, int which, const char* path, int64_t file_offset) {
    CKR(need_map(c));
    const float* src = which == 0 ? c->A : (which == 2 ? c->Out2 : c->Out);
    if (!src) return fail("no output map resident");
    const int fd = open(path, O_RDWR | O_CREAT, 0644);
    if (fd < 0) return fail("gcx_map_download_file: cannot open %s: %s", path, strerror(errno));
    const size_t width = sizeof(float) * (size_t)c->n;
    const int64_t base = file_offset + (int64_t)c->row0 * (int64_t)width;
    const int64_t file_bytes = file_offset + (int64_t)c->n * (int64_t)width;
    struct stat sb;
    if (fstat(fd, &sb) != 0 || (c->world == 1 && sb.st_size != file_bytes && ftruncate(fd, (off_t)file_bytes) != 0) ||
        (c->world > 1 && sb.st_size < file_bytes)) {
        close(fd);
        return fail("gcx_map_download_file: cannot size %s to %lld bytes: %s", path, (long long)file_bytes, strerror(errno));
    }
    char* map = (char*)mmap(nullptr, (size_t)file_bytes, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
    close(fd);
    if (map == MAP_FAILED) return fail("gcx_map_download_file: cannot map %s: %s", path, strerror(errno));
    int rc = ensure_stage(c);
    const int64_t rows_per = std::max<int64_t>(1, (int64_t)(c->stage_bytes / width));
    int64_t r0 = 0, prev_r0 = -1, prev_rows = 0;
    int it = 0;
    while (rc == 0 && (r0 < c->nloc || prev_r0 >= 0)) {
        const int b = it & 1;
        int64_t rows = 0;
        if (r0 < c->nloc) {
            rows = std::min(rows_per, c->nloc - r0);
            if (cudaMemcpy2DAsync(c->stage[b], width, src + r0 * c->ld, sizeof(float) * c->ld, width, (size_t)rows, cudaMemcpyDeviceToHost, c->stream) != cudaSuccess ||
                cudaEventRecord(c->stage_ev[b], c->stream) != cudaSuccess) { rc = fail("gcx_map_download_file: device-to-host copy failed"); break; }
        }
        if (prev_r0 >= 0) {
            if (cudaEventSynchronize(c->stage_ev[b ^ 1]) != cudaSuccess) { rc = fail("cudaEventSynchronize failed"); break; }
            const char* from = (const char*)c->stage[b ^ 1];
            char* to = map + base + prev_r0 * (int64_t)width;
            const int64_t total = prev_rows * (int64_t)width;
            const int64_t gran = 2ll << 20, items = (total + gran - 1) / gran;
            par_for(items, (size_t)total, [&](int64_t a, int64_t e) {
                const int64_t p = a * gran, pe = std::min(total, e * gran);
                if (p < pe) memcpy(to + p, from + p, (size_t)(pe - p));
            });
        }
        if (rows > 0) { prev_r0 = r0; prev_rows = rows; r0 += rows; } else { prev_r0 = -1; }
        ++it;
    }
    munmap(map, (size_t)file_bytes);
    if (rc) cudaStreamSynchronize(c->stream);
    return rc;
}

// Device-resident twin of upload / download (SURVEY.md 8b "when data is already resident"): the map comes from / goes to memory of
// THIS device owned by another producer (a torch tensor, an importer kernel) -- one device-to-device copy, no PCIe trip.
extern "C" int gcx_map_adopt_device(gcx_ctx* c, const float* dev_rows, int64_t dev_ld) {
    CKR(need_map(c));
    CKR(fence_copy(c));
    if (!dev_rows || dev_ld < c->n) return fail("gcx_map_adopt_device: bad pointer / pitch");
    c->sym_state = -1; c->have_rowsum = false; c->have_scale = false;
    CK(cudaMemcpy2DAsync(c->A, sizeof(float) * c->ld, dev_rows, sizeof(float) * dev_ld, sizeof(float) * (size_t)c->n, (size_t)c->nloc,
                         cudaMemcpyDeviceToDevice, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    return 0;
}

extern "C" int gcx_map_export_device(gcx_ctx* c, int which, float* dev_rows, int64_t dev_ld) {
    CKR(need_map(c));
    const float* src = which == 0 ? c->A : (which == 2 ? c->Out2 : c->Out);
    if (!src) return fail("no output map resident");
    if (!dev_rows || dev_ld < c->n) return fail("gcx_map_export_device: bad pointer / pitch");
    CK(cudaMemcpy2DAsync(dev_rows, sizeof(float) * dev_ld, src, sizeof(float) * c->ld, sizeof(float) * (size_t)c->n, (size_t)c->nloc,
                         cudaMemcpyDeviceToDevice, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    return 0;
}

extern "C" int gcx_map_download_async(gcx_ctx* c, int which, float* host, int64_t host_ld) {
    CKR(need_map(c));
    const float* src = which == 0 ? c->A : (which == 2 ? c->Out2 : c->Out);
    if (!src) return fail("no output map resident");
    if (host_ld < c->n) return fail("host_ld < n");
    if (!is_pinned(host)) return fail("gcx_map_download_async needs pinned host memory (gcx_host_alloc)");
    CK(cudaEventRecord(c->ev_ready, c->stream));
    CK(cudaStreamWaitEvent(c->copy_stream, c->ev_ready, 0));
    if (host_ld == c->ld) {     // same pitch on both sides: one contiguous DMA transfer (pad columns travel too)
        CK(cudaMemcpyAsync(host, src, sizeof(float) * (size_t)c->nloc * c->ld, cudaMemcpyDeviceToHost, c->copy_stream));
    } else {
        CK(cudaMemcpy2DAsync(host, sizeof(float) * host_ld, src, sizeof(float) * c->ld, sizeof(float) * (size_t)c->n, (size_t)c->nloc,
                             cudaMemcpyDeviceToHost, c->copy_stream));
    }
    CK(cudaEventRecord(c->ev_copied, c->copy_stream));
    c->copy_pending = true;
    return 0;
}

extern "C" int gcx_download_wait(gcx_ctx* c) {
    CKR(set_dev(c));
    if (c->copy_pending) {
        CK(cudaEventSynchronize(c->ev_copied));
        c->copy_pending = false;
    }
    return 0;
}

// kernels that are about to overwrite a map buffer wait (on the device) for an in-flight async download of it
static int fence_copy(gcx_ctx* c) {
    if (c->copy_pending) CK(cudaStreamWaitEvent(c->stream, c->ev_copied, 0));
    return 0;
}

static int grid_rows(gcx_ctx* c) { return (int)std::min<int64_t>(c->nloc, (int64_t)c->sm_count * 8); }

extern "C" int gcx_map_synth(gcx_ctx* c, uint64_t seed, double scale, double alpha, double empty_frac) {
    CKR(need_map(c));
    CKR(fence_copy(c));
    c->sym_state = -1;
    const unsigned int thresh = (unsigned int)(std::min(1.0, std::max(0.0, empty_frac)) * 16777216.0);
    {
        ProfScope ps(c, 7);
        k_synth<<<grid_rows(c), MV_THREADS, 0, c->stream>>>(c->A, c->ld, (int)c->nloc, (int)c->n, (int)c->row0, seed, (float)scale, (float)alpha, thresh);
    }
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(c->stream));
    c->have_rowsum = false;
    return 0;
}

static int clamp_scale(gcx_ctx* c, int hmin, float vmin, int hmax, float vmax, int hs, float s) {
    CKR(fence_copy(c));
    c->sym_state = -1;
    const long long total4 = (long long)c->nloc * c->ld / 4;
    {
        ProfScope ps(c, 7);
        k_clamp_scale<<<c->sm_count * 8, MV_THREADS, 0, c->stream>>>(c->A, total4, hmin, vmin, hmax, vmax, hs, s);
    }
    CK(cudaGetLastError());
    c->have_rowsum = false;
    return 0;
}

extern "C" int gcx_map_clamp(gcx_ctx* c, int has_vmin, float vmin, int has_vmax, float vmax) {
    CKR(need_map(c));
    if (!has_vmin && !has_vmax) return 0;
    return clamp_scale(c, has_vmin, vmin, has_vmax, vmax, 0, 1.f);
}

extern "C" int gcx_map_scale(gcx_ctx* c, float factor) {
    CKR(need_map(c));
    return clamp_scale(c, 0, 0.f, 0, 0.f, 1, factor);
}

extern "C" int gcx_map_swap(gcx_ctx* c) {
    CKR(need_map(c));
    if (!c->Out) return fail("no output map resident");
    std::swap(c->A, c->Out);
    c->have_rowsum = false;
    c->sym_state = -1;
    return 0;
}

extern "C" int gcx_map_checksum(gcx_ctx* c, int which, uint64_t* sum) {
    CKR(need_map(c));
    const float* src = which == 0 ? c->A : (which == 2 ? c->Out2 : c->Out);
    if (!src) return fail("no output map resident");
    CK(cudaMemsetAsync(c->csum_dev, 0, sizeof(unsigned long long), c->stream));
    k_checksum<<<grid_rows(c), MV_THREADS, 0, c->stream>>>(src, c->ld, (int)c->nloc, (int)c->n, (int)c->row0, c->csum_dev);
    CK(cudaGetLastError());
    c->launches += 1;
    CKR(read_small(c, c->csum_dev, sizeof(unsigned long long), sum));
    return 0;
}

// ---------------------------------------------------------------------------------------------------
// collectives (no-ops on a single rank)
// ---------------------------------------------------------------------------------------------------
__global__ void k_zero_outside_f64(double* v, int nv, int lo, int hi) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nv; i += gridDim.x * blockDim.x)
        if (i < lo || i >= hi) v[i] = 0.0;
}
__global__ void k_zero_outside_i32(int* v, int nv, int lo, int hi) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nv; i += gridDim.x * blockDim.x)
        if (i < lo || i >= hi) v[i] = 0;
}

static int allreduce_f64(gcx_ctx* c, double* vec, size_t count) {
    if (c->world == 1) return 0;
    NK(g_nccl.AllReduce(vec, vec, count, ncclFloat64, ncclSum, c->comm, c->stream));
    return 0;
}

// Every rank wrote its own rows [row0, row0+nloc) of an n-vector; afterwards every rank holds all of it.
// Equal row blocks: one in-place NCCL allgather.  Custom (unequal) blocks: zero the foreign entries and allreduce
// (x + 0 + ... + 0 is exact, so the result is the same bits on every rank).
static int allgather_f64(gcx_ctx* c, double* vec) {
    if (c->world == 1) return 0;
    if (c->exchange_allreduce) {
        k_zero_outside_f64<<<64, 256, 0, c->stream>>>(vec, (int)c->nv, (int)c->row0, (int)(c->row0 + c->nloc));
        CK(cudaGetLastError());
        return allreduce_f64(c, vec, (size_t)c->n);
    }
    NK(g_nccl.AllGather(vec + c->rank * c->rpr, vec, (size_t)c->rpr, ncclFloat64, c->comm, c->stream));
    return 0;
}
static int allgather_i32(gcx_ctx* c, int32_t* vec) {
    if (c->world == 1) return 0;
    if (c->exchange_allreduce) {
        k_zero_outside_i32<<<64, 256, 0, c->stream>>>(vec, (int)c->nv, (int)c->row0, (int)(c->row0 + c->nloc));
        CK(cudaGetLastError());
        NK(g_nccl.AllReduce(vec, vec, (size_t)c->n, ncclInt32, ncclSum, c->comm, c->stream));
        return 0;
    }
    NK(g_nccl.AllGather(vec + c->rank * c->rpr, vec, (size_t)c->rpr, ncclInt32, c->comm, c->stream));
    return 0;
}

// Peer-mapped exchange buffers of the sharded symmetric IC pass.  Collective: every rank calls it (from gcx_map_create).  Each rank
// allocates [world][G] flags + [2][world][nv] doubles, the ranks swap {cudaIpc handle, pid, pointer, device, G, nv} with one NCCL
// allgather, and map each other's buffer: cudaIpcOpenMemHandle across processes, plain peer access inside one process (one host
// thread per GPU).  Any failure leaves xchg_ok = false on EVERY rank (the verdicts are allreduced) and the pass keeps its NCCL exchange.
struct XchgBlob {
    cudaIpcMemHandle_t handle;     // 64 bytes
    long long pid;
    void* ptr;
    int device, grid;
    long long nv;
    int ok, pad;
};
static_assert(sizeof(XchgBlob) <= 128, "blob must fit its allgather slot");

static size_t xchg_flag_bytes(gcx_ctx* c) { return (sizeof(unsigned int) * (size_t)c->world * c->sym_grid + 255) / 256 * 256; }

static int xchg_setup(gcx_ctx* c) {
    xchg_free(c);
    if (c->world < 2 || c->world > XCHG_MAX_WORLD || !c->comm || c->sym_grid < 1) return 0;
    const int W = c->world;
    const size_t fb = xchg_flag_bytes(c), bytes = fb + sizeof(double) * 2 * (size_t)W * c->nv;
    unsigned char* stage = nullptr;       // device staging of the W blobs
    CK(cudaMalloc(&stage, 128 * (size_t)W));
    XchgBlob mine;
    memset(&mine, 0, sizeof mine);
    bool good = cudaMalloc(&c->xchg_buf, bytes) == cudaSuccess;
    if (good) good = cudaMemsetAsync(c->xchg_buf, 0, bytes, c->stream) == cudaSuccess;
    if (good) good = cudaIpcGetMemHandle(&mine.handle, c->xchg_buf) == cudaSuccess;
    cudaGetLastError();
    mine.pid = (long long)getpid();
    mine.ptr = c->xchg_buf;
    mine.device = c->device;
    mine.grid = c->sym_grid;
    mine.nv = c->nv;
    mine.ok = good ? 1 : 0;
    std::vector<unsigned char> host(128 * (size_t)W, 0);
    memcpy(host.data() + 128 * (size_t)c->rank, &mine, sizeof mine);
    CK(cudaMemcpyAsync(stage + 128 * (size_t)c->rank, host.data() + 128 * (size_t)c->rank, 128, cudaMemcpyHostToDevice, c->stream));
    NK(g_nccl.AllGather(stage + 128 * (size_t)c->rank, stage, 128, ncclUint8, c->comm, c->stream));
    CK(cudaMemcpyAsync(host.data(), stage, 128 * (size_t)W, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    for (int q = 0; q < W && good; ++q) {
        XchgBlob b;
        memcpy(&b, host.data() + 128 * (size_t)q, sizeof b);
        if (!b.ok || b.grid != c->sym_grid || b.nv != c->nv) { good = false; break; }
        if (q == c->rank) { c->xchg_peer[q] = c->xchg_buf; continue; }
        if (b.pid == mine.pid) {
            int can = 0;
            if (cudaDeviceCanAccessPeer(&can, c->device, b.device) != cudaSuccess || !can) { good = false; break; }
            cudaError_t e = cudaDeviceEnablePeerAccess(b.device, 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) { good = false; break; }
            cudaGetLastError();
            c->xchg_peer[q] = b.ptr;
        } else {
            void* p = nullptr;
            if (cudaIpcOpenMemHandle(&p, b.handle, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { good = false; break; }
            c->xchg_peer[q] = p;
            c->xchg_ipc[q] = true;
        }
    }
    cudaGetLastError();
    // every rank must take the same path: min over ranks of the local verdict (device int, reusing the staging buffer)
    int verdict = good ? 1 : 0;
    CK(cudaMemcpyAsync(stage, &verdict, sizeof(int), cudaMemcpyHostToDevice, c->stream));
    NK(g_nccl.AllReduce(stage, stage, 1, ncclInt32, ncclMin, c->comm, c->stream));
    CK(cudaMemcpyAsync(&verdict, stage, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    CK(cudaFree(stage));
    if (!verdict) { xchg_free(c); return 0; }
    c->xchg_ok = true;
    return 0;
}

static IcXchg xchg_args(gcx_ctx* c) {
    IcXchg X;
    memset(&X, 0, sizeof X);
    X.world = c->world; X.rank = c->rank; X.seq0 = c->xchg_seq;
    const size_t fb = xchg_flag_bytes(c);
    X.flags = (unsigned int*)c->xchg_buf;
    X.recv = (const double*)((char*)c->xchg_buf + fb);
    for (int q = 0; q < c->world && q < XCHG_MAX_WORLD; ++q) {
        X.peer_flags[q] = (unsigned int*)c->xchg_peer[q];
        X.peer_recv[q] = (double*)((char*)c->xchg_peer[q] + fb);
    }
    return X;
}

extern "C" int gcx_set_option(gcx_ctx* c, const char* key, double value) {
    if (!c || !key) return fail("gcx_set_option: null argument");
    const std::string k(key);
    if (k == "sym") c->sym_enable = (int)value;
    else if (k == "assume_symmetric") { c->assume_symmetric = (int)value; c->sym_state = -1; }
    else if (k == "exchange_allreduce") c->exchange_allreduce = (int)value;
    else if (k == "peer_exchange") c->xchg_enable = (int)value;          // before gcx_map_create, on every rank
    else if (k == "ic_defer") c->ic_defer = (int)value;
    else if (k == "ic_persist") c->ic_persist = (int)value;
    else if (k == "sym_blocks") {
        if (c->sym_blocks != (int)value && c->stream) {       // the unit tables depend on it: rebuild on the next run
            cudaSetDevice(c->device);
            cudaStreamSynchronize(c->stream);
            free_sym(c);
        }
        c->sym_blocks = (int)value;
    }
    else if (k == "sym_balance") {       // before the first solve, on every rank (the tables are built once)
        if (c->sym_balance != (int)value && c->stream) {
            cudaSetDevice(c->device);
            cudaStreamSynchronize(c->stream);
            free_sym(c);
        }
        c->sym_balance = (int)value;
    }
    else if (k == "matvec_variant") c->mv_variant = (int)value;
    else if (k == "sd_variant") c->sd_variant = (int)value;
    else if (k == "corr_variant") c->corr_variant = (int)value;
    else if (k == "corr_int_path") c->corr_int_path = (int)value;
    else if (k == "corr_own_gemm") c->corr_own_gemm = (int)value;
    else return fail("gcx_set_option: unknown key '%s'", key);
    return 0;
}

extern "C" int gcx_dist_unique_id(void* id128) {
    CKR(nccl_load());
    ncclUniqueId id;
    NK(g_nccl.GetUniqueId(&id));
    memcpy(id128, &id, 128);
    return 0;
}

extern "C" int gcx_dist_init(gcx_ctx* c, int rank, int world, const void* id128) {
    CKR(set_dev(c));
    if (c->A) return fail("gcx_dist_init must precede gcx_map_create");
    if (c->comm) return fail("gcx_dist_init: this context already has a communicator");
    if (!id128 && world > 1) return fail("gcx_dist_init: null unique id");
    if (world < 1 || rank < 0 || rank >= world) return fail("bad rank/world");
    c->rank = rank; c->world = world;
    if (world == 1) return 0;
    CKR(nccl_load());
    ncclUniqueId id;
    memcpy(&id, id128, 128);
    NK(g_nccl.CommInitRank(&c->comm, world, id, rank));
    return 0;
}

// ---------------------------------------------------------------------------------------------------
// mask
// ---------------------------------------------------------------------------------------------------
static int run_row_stats(gcx_ctx* c) {
    if (c->have_rowsum) return 0;
    {
        ProfScope ps(c, 3);
        k_row_stats<<<grid_rows(c), MV_THREADS, 0, c->stream>>>(c->A, c->ld, (int)c->nloc, (int)c->n, c->vrowsum + c->row0, c->zero_count + c->row0,
                                                                  c->neg_flag + c->row0);
    }
    CK(cudaGetLastError());
    CKR(allgather_f64(c, c->vrowsum));
    CKR(allgather_i32(c, c->zero_count));
    CKR(allgather_i32(c, c->neg_flag));
    c->have_rowsum = true;
    return 0;
}

/* 1 per row that holds negative entries (n int32; computed with gcx_row_stats): lib/ccmapHelpers.pyx:192 tests emptiness with a
 * float32 sum, which only differs from "all entries zero" on such rows -- the caller re-checks them in NumPy's float32 arithmetic */
extern "C" int gcx_row_negative_flags(gcx_ctx* c, int32_t* flags) {
    CKR(need_map(c));
    CKR(run_row_stats(c));
    if (!flags) return fail("gcx_row_negative_flags: null output");
    CK(cudaMemcpyAsync(flags, c->neg_flag, sizeof(int32_t) * c->n, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    return 0;
}

extern "C" int gcx_row_stats(gcx_ctx* c, double* rowsum, int32_t* zero_count) {
    CKR(need_map(c));
    CKR(run_row_stats(c));
    if (rowsum) CK(cudaMemcpyAsync(rowsum, c->vrowsum, sizeof(double) * c->n, cudaMemcpyDeviceToHost, c->stream));
    if (zero_count) CK(cudaMemcpyAsync(zero_count, c->zero_count, sizeof(int32_t) * c->n, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    return 0;
}

extern "C" int gcx_set_mask(gcx_ctx* c, const uint8_t* keep) {
    CKR(need_map(c));
    if (!keep) return fail("gcx_set_mask: null mask");
    CK(cudaMemsetAsync(c->keep, 0, (size_t)c->nv, c->stream));
    CK(cudaMemcpyAsync(c->keep, keep, (size_t)c->n, cudaMemcpyHostToDevice, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    c->have_mask = true;
    c->have_scale = false;
    return 0;
}

// ---------------------------------------------------------------------------------------------------
// matvec dispatch
// ---------------------------------------------------------------------------------------------------
// ---------------------------------------------------------------------------------------------------
// symmetric (upper-triangle) passes
// ---------------------------------------------------------------------------------------------------
static int64_t sym_bytes_per_pass(gcx_ctx* c) { return c->sym_bytes; }

// The aligned row partition: cuts at r*n/world rounded to 1024 rows (the panel width of the symmetric pass).
static int64_t aligned_cut(int64_t n, int world, int r) {
    if (r <= 0) return 0;
    if (r >= world) return n;
    const int64_t x = (2 * (int64_t)r * n / world + 1024) / 2048 * 1024;
    return std::min(n, x);
}

extern "C" int gcx_partition_rows(int64_t n, int world, int rank, int64_t* row0, int64_t* nloc) {
    if (world < 1 || rank < 0 || rank >= world || n < 1) return fail("gcx_partition_rows: bad arguments");
    *row0 = aligned_cut(n, world, rank);
    *nloc = aligned_cut(n, world, rank + 1) - *row0;
    return 0;
}

// ---- which rows of which column panels a rank walks in the upper-triangle pass (pure host logic, shared by the table builder
// and by gcx_debug_sym_plan, which the CPU tests use to check that every unordered pair of bins is visited exactly once) ----
//
// Sharded + aligned partition: rank r owns the triangle of its diagonal block; every off-diagonal block pair {lo, hi} (lo < hi)
// is shared by its two ranks through a panel-aligned cut inside lo's rows: rank lo walks rows [a_lo, cut) of block (lo, hi) from
// its own rows, rank hi walks ALL its rows against the columns [cut, b_lo) of the mirrored block (hi, lo).  cut = b_lo gives the
// whole pair to lo, cut = a_lo to hi.  The start is the circulant deal (lo takes the pair if hi - lo < W/2, hi if > W/2, the
// antipodal pair is halved); the cuts are then moved panel by panel from the most loaded rank to a partner while that lowers
// the maximum -- the aligned row blocks differ by a panel (8 vs 9 at chr1-like sizes on 8 GPUs: +-5 % of a pass, which every
// other rank spent waiting in the exchange, profiles/r02_stamps_g_8gpu_final.txt).
struct SymPlan {
    std::vector<long long> nb;     // bands of SYM_R rows walked in panel k (from the rank's first row)
    std::vector<int> full;         // 1: off-diagonal block (no triangle mask)
    int64_t bytes = 0;
};

static bool sym_blocks_possible(int64_t n, int W) {
    for (int r = 0; r < W; ++r)
        if (aligned_cut(n, W, r + 1) - aligned_cut(n, W, r) < 2048) return false;
    return W > 1;
}

static void sym_panel_plan(int64_t n, int64_t ld, int W, int me, int64_t r0, int64_t r1, bool blocks, const std::vector<int64_t>& cut, SymPlan* P) {
    const int nch = (int)((ld / 4 + MV_THREADS - 1) / MV_THREADS);
    const long long all_bands = (r1 - r0 + SYM_R - 1) / SYM_R;
    P->nb.assign((size_t)nch, 0);
    P->full.assign((size_t)nch, 0);
    P->bytes = 0;
    for (int k = 0; k < nch; ++k) {
        long long nb = 0;
        int full = 0;
        const long long col0 = (long long)k * 1024;
        int s_blk = 0;
        if (blocks) {
            while (s_blk + 1 < W && aligned_cut(n, W, s_blk + 1) <= col0) ++s_blk;
        }
        if (!blocks || s_blk == me) {
            // own diagonal region (or the whole upper triangle of the rank's rows when blocks are off)
            const long long rows_end = std::min<long long>(r1, (k < nch - 1) ? (long long)(k + 1) * 1024 : (long long)n);
            nb = rows_end > r0 ? (rows_end - r0 + SYM_R - 1) / SYM_R : 0;
            if (blocks && col0 >= r1) nb = 0;
        } else {
            full = 1;
            const int lo = std::min(me, s_blk), hi = std::max(me, s_blk);
            const int64_t ct = cut[(size_t)lo * W + hi];
            if (me == lo) nb = (ct - r0) / SYM_R;                          // first rows of the block, up to the cut
            else nb = (col0 >= ct) ? all_bands : 0;                       // all rows against the partner's columns from the cut on
        }
        P->full[k] = full;
        P->nb[k] = nb;
        P->bytes += nb * SYM_R * std::min<int64_t>(1024, ld - (int64_t)k * 1024) * 4;
    }
}

static void sym_pair_cuts(int64_t n, int64_t ld, int W, bool balance, std::vector<int64_t>* cut) {
    cut->assign((size_t)W * W, 0);
    for (int lo = 0; lo < W; ++lo)
        for (int hi = lo + 1; hi < W; ++hi) {
            const int64_t a = aligned_cut(n, W, lo), b = aligned_cut(n, W, lo + 1);
            const int d = hi - lo;
            (*cut)[(size_t)lo * W + hi] = 2 * d < W ? b : (2 * d > W ? a : a + ((b - a) / 2 + 512) / 1024 * 1024);
        }
    if (!balance || W < 3) return;
    std::vector<int64_t> work((size_t)W);
    SymPlan P;
    auto measure = [&]() {
        for (int r = 0; r < W; ++r) {
            sym_panel_plan(n, ld, W, r, aligned_cut(n, W, r), aligned_cut(n, W, r + 1), true, *cut, &P);
            work[r] = P.bytes;
        }
    };
    // steepest descent on (max, sum of squares): try every one-panel move of every pair, keep the best, stop when none helps
    auto score = [&](int64_t* mx, double* sq) {
        measure();
        *mx = *std::max_element(work.begin(), work.end());
        *sq = 0.0;
        for (int r = 0; r < W; ++r) *sq += (double)work[r] * (double)work[r];
    };
    int64_t cur_max;
    double cur_sq;
    score(&cur_max, &cur_sq);
    for (int iter = 0; iter < 16 * W * W; ++iter) {
        int best_lo = -1, best_hi = -1;
        int64_t best_step = 0, best_max = cur_max;
        double best_sq = cur_sq;
        for (int lo = 0; lo < W; ++lo)
            for (int hi = lo + 1; hi < W; ++hi)
                for (int64_t step = -1024; step <= 1024; step += 2048) {
                    int64_t& ct = (*cut)[(size_t)lo * W + hi];
                    const int64_t a = aligned_cut(n, W, lo), b = aligned_cut(n, W, lo + 1);
                    if (ct + step < a || ct + step > b) continue;
                    ct += step;
                    int64_t m;
                    double q;
                    score(&m, &q);
                    ct -= step;
                    if (m < best_max || (m == best_max && q < best_sq)) { best_max = m; best_sq = q; best_lo = lo; best_hi = hi; best_step = step; }
                }
        if (best_lo < 0) break;
        (*cut)[(size_t)best_lo * W + best_hi] += best_step;
        cur_max = best_max;
        cur_sq = best_sq;
    }
}

/* Testing hook (no GPU needed): the panel plan of `rank` for a map of n bins row-sharded over `world` ranks with the library's
   aligned partition -- nb[k] bands of 4 rows from the rank's first row in column panel k (1024 columns), full[k] = 1 for an
   off-diagonal block.  nch = number of panels = ceil(round_up(n, 32) / 1024); both arrays hold nch entries. */
extern "C" int gcx_debug_sym_plan(int64_t n, int world, int rank, int balance, int64_t* nb, int32_t* full, int nch, int64_t* bytes) {
    if (n < 1 || world < 1 || rank < 0 || rank >= world || !nb || !full) return fail("gcx_debug_sym_plan: bad arguments");
    const int64_t ld = round_up(n, 32);
    const int want = (int)((ld / 4 + MV_THREADS - 1) / MV_THREADS);
    if (nch != want) return fail("gcx_debug_sym_plan: a map of %lld bins has %d panels", (long long)n, want);
    const bool blocks = sym_blocks_possible(n, world);
    std::vector<int64_t> cut;
    if (blocks) sym_pair_cuts(n, ld, world, balance != 0, &cut);
    SymPlan P;
    sym_panel_plan(n, ld, world, rank, aligned_cut(n, world, rank), aligned_cut(n, world, rank + 1), blocks, cut, &P);
    for (int k = 0; k < nch; ++k) { nb[k] = P.nb[k]; full[k] = P.full[k]; }
    if (bytes) *bytes = P.bytes;
    return 0;
}

// decides (and caches) whether the resident map can use the upper-triangle passes, and tabulates the geometry
static int sym_ready(gcx_ctx* c, bool* ok) {
    *ok = false;
    if (!c->sym_enable || c->n < 2048) return 0;
    // sharded: needs the zero-fill + allreduce exchange (an option every rank sets) and panel-aligned cuts; a rank whose block is
    // not aligned cannot simply return -- the others would wait in the collective of the symmetry test -- it votes "no" there
    if (c->world > 1 && !c->exchange_allreduce) return 0;
    const bool aligned = c->world == 1 || c->row0 % 1024 == 0;
    if (c->world > 1 && c->assume_symmetric && !aligned) return 0;
    if (c->world == 1 && (c->row0 != 0 || c->nloc != c->n)) return 0;
    if (c->sym_state < 0) {
        if (c->assume_symmetric) {
            c->sym_state = 1;             // the caller vouches (gcx_set_option "assume_symmetric"): no test
        } else if (c->world > 1) {
            // sharded: hash sums of the entries above / below the diagonal, added over the ranks, must agree (k_sym_hash)
            unsigned long long* hs = nullptr;
            CKR(ensure_scratch(c, 256));
            hs = (unsigned long long*)c->scratch;
            CK(cudaMemsetAsync(hs, 0, 2 * sizeof(unsigned long long), c->stream));
            k_sym_hash<<<grid_rows(c), MV_THREADS, 0, c->stream>>>(c->A, c->ld, (int)c->nloc, (int)c->n, (int)c->row0, hs);
            CK(cudaGetLastError());
            c->launches += 1;
            if (!aligned) CK(cudaMemsetAsync(hs, 0xff, sizeof(unsigned long long), c->stream));      // poison: the sums cannot agree
            NK(g_nccl.AllReduce(hs, hs, 2, ncclUint64, ncclSum, c->comm, c->stream));
            unsigned long long h2[2] = {0, 1};
            CKR(read_small(c, hs, sizeof h2, h2));
            c->sym_state = h2[0] == h2[1] ? 1 : 0;
        } else {
            CK(cudaMemsetAsync(c->csum_dev, 0, sizeof(unsigned long long), c->stream));
            k_count_asym<<<c->sm_count * 8, 256, 0, c->stream>>>(c->A, c->ld, (int)c->n, c->csum_dev);
            CK(cudaGetLastError());
            c->launches += 1;
            unsigned long long bad = 0;
            CKR(read_small(c, c->csum_dev, sizeof bad, &bad));
            c->sym_state = bad == 0 ? 1 : 0;
        }
    }
    if (c->sym_state != 1) return 0;
    if (!c->rowpart) {
        free_sym(c);          // nothing stale survives an earlier, failed build
        // every panel's and every CTA's share of the panel-major unit list (the kernels do no 64-bit divisions)
        const int n = (int)c->n, n4 = (int)(c->ld / 4), nch = (n4 + MV_THREADS - 1) / MV_THREADS, G = c->sym_grid;
        const long long r0 = c->row0, r1 = c->row0 + c->nloc;
        std::vector<long long> pu((size_t)nch + 1, 0);
        std::vector<int> pfull((size_t)nch, 0);
        const int W = c->world, me = c->rank;
        // sharded + aligned partition: the diagonal block plus a share of every off-diagonal block pair (sym_panel_plan above)
        bool blocks = W > 1 && c->sym_blocks && sym_blocks_possible(n, W);
        if (blocks) blocks = (r0 == aligned_cut(n, W, me) && r1 == aligned_cut(n, W, me + 1));
        std::vector<int64_t> cuts;
        if (blocks) sym_pair_cuts(n, c->ld, W, c->sym_balance != 0, &cuts);
        SymPlan plan;
        sym_panel_plan(n, c->ld, W, me, r0, r1, blocks, cuts, &plan);
        const int64_t bytes = plan.bytes;
        int kfold_lo = -1;
        std::vector<long long> nbk(plan.nb);
        for (int k = 0; k < nch; ++k) {
            pfull[k] = plan.full[k];
            if (nbk[k] > 0 && kfold_lo < 0) kfold_lo = k;
        }
        c->sym_kfold_lo = blocks ? kfold_lo : -1;
        // dynamic pool: ~GCX_SYM_DYN percent of the units, cut into chunks of DYN_CHUNK consecutive bands of one panel.
        // GCX_SYM_POOL=1 (default): the LAST bands of every large full-width panel -- a bin's fold then adds a handful of chunk
        // partials whichever panel it lies in; =0: whole panels, largest first (round 1: the ~2000 bins of the pooled panels each
        // folded ~400 chunk partials and held the whole grid at the barrier, profiles/r02_phase_stamps.md).
        // defaults from the sweeps in profiles/r02_ic_variants_*.txt: 30 % of the units in guided chunks of 64 / 32 / 16 / 8 bands at
        // chr1 @ 10 kb and above; smaller maps (fewer units per CTA) do better with 20 % in 32 / 16 / 8 / 4
        int DYN_CHUNK = c->n >= 20000 ? 64 : 32;
        if (const char* e = getenv("GCX_SYM_DYN_CHUNK")) DYN_CHUNK = std::max(4, atoi(e));
        double dyn_frac = c->n >= 20000 ? 0.30 : 0.20;
        if (const char* e = getenv("GCX_SYM_DYN")) dyn_frac = atof(e) / 100.0;
        int pool_mode = 1;
        if (const char* e = getenv("GCX_SYM_POOL")) pool_mode = atoi(e);
        long long Uall = 0;
        for (int k = 0; k < nch; ++k) Uall += nbk[k];
        std::vector<long long> pool_bands((size_t)nch, 0);       // bands at the END of panel k that belong to the pool
        if (pool_mode == 0) {
            std::vector<int> order;
            for (int k = 0; k < nch; ++k)
                if (nbk[k] > 0 && std::min<int64_t>(1024, c->ld - (int64_t)k * 1024) == 1024) order.push_back(k);
            std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return nbk[a] != nbk[b] ? nbk[a] > nbk[b] : a > b; });
            long long pooled = 0;
            for (int k : order) {
                if ((double)pooled >= dyn_frac * (double)Uall) break;
                if (Uall - pooled - nbk[k] < 4LL * G) break;             // keep a meaningful static share
                pool_bands[k] = nbk[k];
                pooled += nbk[k];
            }
        } else {
            long long pooled = 0;
            for (int k = nch - 1; k >= 0; --k) {
                if (nbk[k] < 8LL * DYN_CHUNK || std::min<int64_t>(1024, c->ld - (int64_t)k * 1024) != 1024) continue;
                const long long pb = (long long)(dyn_frac * (double)nbk[k] / DYN_CHUNK + 0.5) * DYN_CHUNK;
                if (pb <= 0 || Uall - pooled - pb < 4LL * G) continue;
                pool_bands[k] = std::min(pb, nbk[k]);
                pooled += pool_bands[k];
            }
        }
        // Guided chunk sizes: most of the pool goes out in DYN_CHUNK-band chunks, the last part in chunks of a half, a quarter and an
        // eighth of that, handed out LAST -- the chunk in flight when the pool runs dry bounds every CTA's idle time at the end of a
        // pass (a 32-band chunk is ~23 us of one CTA's bandwidth; the per-CTA throughput varies +-15 % at random).
        // GCX_SYM_GUIDED=0: equal chunks (round 1).
        int guided = 1;
        if (const char* e = getenv("GCX_SYM_GUIDED")) guided = atoi(e);
        std::vector<int> dk, db0, dcnt, dorder, pdyn0((size_t)nch + 1, 0);
        std::vector<std::vector<int>> by_class(4);
        for (int k = 0; k < nch; ++k) {
            pdyn0[k] = (int)dk.size();
            const long long stat = nbk[k] - pool_bands[k];
            long long b = stat;
            const long long pb = pool_bands[k];
            // shares of this panel's pool per size class (bands): 60 % full chunks, then 22 % / 12 % / 6 % in halves
            const double share[4] = {guided ? 0.60 : 1.0, 0.22, 0.12, 0.06};
            long long class_end[4];
            double acc = 0.0;
            for (int cl = 0; cl < 4; ++cl) {
                acc += share[cl];
                class_end[cl] = cl == 3 || !guided ? nbk[k] : stat + (long long)(acc * (double)pb) / DYN_CHUNK * DYN_CHUNK;
            }
            for (int cl = 0; cl < 4 && b < nbk[k]; ++cl) {
                const int size = std::max(4, DYN_CHUNK >> cl);
                while (b < class_end[cl]) {
                    const int cnt = (int)std::min<long long>((long long)size, class_end[cl] - b);
                    by_class[cl].push_back((int)dk.size());
                    dk.push_back(k);
                    db0.push_back((int)b);
                    dcnt.push_back(cnt);
                    b += cnt;
                }
            }
            pu[k + 1] = pu[k] + stat;
        }
        pdyn0[nch] = (int)dk.size();
        for (int cl = 0; cl < 4; ++cl) dorder.insert(dorder.end(), by_class[cl].begin(), by_class[cl].end());     // hand-out order: large first
        const int ndyn = (int)dk.size();
        const long long U = pu[nch];
        if (U < G) return 0;          // tiny share: keep the dense pass
        auto panel_of = [&](long long u) { return (int)(std::upper_bound(pu.begin(), pu.end(), u) - pu.begin()) - 1; };
        std::vector<long long> u0((size_t)G + 1);
        std::vector<int> kf((size_t)G, 0), kl((size_t)G, 0), pc((size_t)nch, G), ph((size_t)nch, -1);
        // static split of the unit list over the CTAs.  The second CTA of every SM (index >= SM count) measurably gets a
        // smaller share of the SM's bandwidth (tools/sym_cta_cycles.py), so it is given `skew` percent fewer units.
        {
            double skew = 0.0;
            if (const char* e = getenv("GCX_SYM_SKEW")) skew = atof(e) / 100.0;
            const int first = std::min(G, c->sm_count);
            const double wa = 1.0 + skew, wb = 1.0 - skew;
            const double W = wa * first + wb * (G - first);
            double acc = 0.0;
            u0[0] = 0;
            for (int g = 0; g < G; ++g) {
                acc += (g < first ? wa : wb);
                u0[g + 1] = (g + 1 == G) ? U : (long long)((double)U * acc / W);
            }
        }
        int slots = 1;
        for (int g = 0; g < G; ++g) {
            if (u0[g + 1] <= u0[g]) continue;
            kf[g] = panel_of(u0[g]);
            kl[g] = panel_of(u0[g + 1] - 1);
            slots = std::max(slots, kl[g] - kf[g] + 1);
            for (int k = kf[g]; k <= kl[g]; ++k) { pc[k] = std::min(pc[k], g); ph[k] = std::max(ph[k], g); }
        }
        c->sym_slots = slots;
        c->sym_nch = nch;
        c->sym_bytes = bytes;
        std::vector<long long> tll(pu);
        tll.insert(tll.end(), u0.begin(), u0.end());
        std::vector<int> ti(kf);
        ti.insert(ti.end(), kl.begin(), kl.end());
        ti.insert(ti.end(), pc.begin(), pc.end());
        ti.insert(ti.end(), ph.begin(), ph.end());
        ti.insert(ti.end(), pfull.begin(), pfull.end());
        ti.insert(ti.end(), dk.begin(), dk.end());
        ti.insert(ti.end(), db0.begin(), db0.end());
        ti.insert(ti.end(), dcnt.begin(), dcnt.end());
        ti.insert(ti.end(), pdyn0.begin(), pdyn0.end());
        ti.insert(ti.end(), dorder.begin(), dorder.end());
        // sharded: the row partials of a bin can only be non-zero in the panels this rank walks (about W/2 + 1 of W column blocks)
        std::vector<int> foldk;
        if (blocks)
            for (int k = 0; k < nch; ++k)
                if (nbk[k] > 0) foldk.push_back(k);
        ti.insert(ti.end(), foldk.begin(), foldk.end());
        c->sym_nfoldk = (int)foldk.size();
        c->sym_ndyn = ndyn;
        CK(cudaMalloc(&c->sym_tab_ll, sizeof(long long) * tll.size()));
        CK(cudaMalloc(&c->sym_tab_i, sizeof(int) * ti.size()));
        CK(cudaMemcpy(c->sym_tab_ll, tll.data(), sizeof(long long) * tll.size(), cudaMemcpyHostToDevice));
        CK(cudaMemcpy(c->sym_tab_i, ti.data(), sizeof(int) * ti.size(), cudaMemcpyHostToDevice));
        const size_t rp_bytes = sizeof(double) * (size_t)nch * c->ld;
        const size_t cp_bytes = sizeof(double) * (size_t)G * slots * SYM_NG * 1024;
        CK(cudaMalloc(&c->rowpart, rp_bytes));
        CK(cudaMalloc(&c->colpart, cp_bytes));
        CK(cudaMalloc(&c->sym_dyn_counter, sizeof(unsigned int)));
        CK(cudaMemsetAsync(c->sym_dyn_counter, 0, sizeof(unsigned int), c->stream));
        if (ndyn > 0) {
            CK(cudaMalloc(&c->colpart_dyn, sizeof(double) * (size_t)ndyn * SYM_NG * 1024));
            CK(cudaMemsetAsync(c->colpart_dyn, 0, sizeof(double) * (size_t)ndyn * SYM_NG * 1024, c->stream));
        }
        if (getenv("GCX_SYM_DEBUG") && !c->sym_dbg) {
            CK(cudaMalloc(&c->sym_dbg, sizeof(long long) * (size_t)G));
            CK(cudaMemset(c->sym_dbg, 0, sizeof(long long) * (size_t)G));
            CK(cudaMalloc(&c->sym_stamps, sizeof(long long) * 16 * (size_t)G));
            CK(cudaMemset(c->sym_stamps, 0, sizeof(long long) * 16 * (size_t)G));
            if (const char* e = getenv("GCX_SYM_DEBUG_LAUNCH")) c->sym_dbg_launch = atoi(e);
        }
        // rowpart must start from zero: (panel, row) pairs this rank never walks are read by the fold
        CK(cudaMemsetAsync(c->rowpart, 0, rp_bytes, c->stream));
        CK(cudaMemsetAsync(c->colpart, 0, cp_bytes, c->stream));
        if (c->world > 1 && c->xchg_ok) {
            // the sharded pass does not send the bins a rank never contributes to: the receive slots must hold zeros for THESE
            // tables (they may differ from an earlier build's).  All ranks build together; the collective keeps any rank from
            // starting a pass -- and storing into a peer -- before that peer's slots are cleared.
            const size_t fb = xchg_flag_bytes(c);
            CK(cudaMemsetAsync((char*)c->xchg_buf + fb, 0, sizeof(double) * 2 * (size_t)W * c->nv, c->stream));
            CKR(ensure_scratch(c, 256));
            CK(cudaMemsetAsync(c->scratch, 0, sizeof(int), c->stream));
            NK(g_nccl.AllReduce(c->scratch, c->scratch, 1, ncclInt32, ncclSum, c->comm, c->stream));
        }
    }
    *ok = true;
    return 0;
}

static SymTab sym_tab(gcx_ctx* c) {
    const int nch = c->sym_nch, G = c->sym_grid;
    SymTab T;
    T.panel_u0 = c->sym_tab_ll;
    T.cta_u0 = c->sym_tab_ll + nch + 1;
    T.cta_kf = c->sym_tab_i;
    T.cta_kl = c->sym_tab_i + G;
    T.panel_cta = c->sym_tab_i + 2 * G;
    T.panel_cta_hi = c->sym_tab_i + 2 * G + nch;
    T.panel_full = c->sym_tab_i + 2 * G + 2 * nch;
    T.kfold_lo = c->sym_kfold_lo;
    T.dbg_cycles = c->sym_dbg;
    T.dbg_stamps = c->sym_stamps;
    T.dbg_launch = c->sym_dbg_launch;
    T.l2_hint = c->sym_l2_hint;
    T.ndyn = c->sym_ndyn;
    T.dyn_k = c->sym_tab_i + 2 * G + 3 * nch;
    T.dyn_b0 = T.dyn_k + c->sym_ndyn;
    T.dyn_cnt = T.dyn_b0 + c->sym_ndyn;
    T.panel_dyn0 = T.dyn_cnt + c->sym_ndyn;
    T.dyn_order = T.panel_dyn0 + nch + 1;
    T.fold_k = T.dyn_order + c->sym_ndyn;
    T.n_fold_k = c->sym_nfoldk;
    T.colpart_dyn = c->colpart_dyn;
    T.dyn_counter = c->sym_dyn_counter;
    T.nslots = c->sym_slots;
    T.row0 = (int)c->row0;
    T.nloc = (int)c->nloc;
    return T;
}

// q = A z through the upper-triangle pass; sharded: every rank folds its partial t, an allreduce adds them
static int matvec_sym(gcx_ctx* c, const double* z, double* q, const int* done) {
    const size_t smem = (size_t)SYM_STAGES * SYM_R * MV_THREADS * sizeof(float4);
    const int n = (int)c->n, n4 = (int)(c->ld / 4);
    if (c->sym_ndyn > 0) CK(cudaMemsetAsync(c->sym_dyn_counter, 0, sizeof(unsigned int), c->stream));
    {
        ProfScope ps(c, 0);
        k_matvec_sym_tma<SYM_STAGES, SYM_OCC><<<c->sym_grid, SYM_THREADS, smem, c->stream>>>(c->A, c->ld, n4, z, c->rowpart, c->colpart, sym_tab(c), done);
        k_sym_fold<<<(n * 8 + 255) / 256, 256, 0, c->stream>>>(n, n4, c->rowpart, c->colpart, sym_tab(c), q, done);
    }
    c->launches += 1;
    CK(cudaGetLastError());
    if (c->world > 1) CKR(allreduce_f64(c, q, (size_t)c->n));
    return 0;
}

template <int R, int OCC>
static void launch_mv(gcx_ctx* c, const double* z, double* q, const int* done) {
    k_matvec<R, OCC><<<c->sm_count * OCC, MV_THREADS, 0, c->stream>>>(c->A, c->ld, (int)c->nloc, (int)(c->ld / 4), z, q + c->row0, c->ws, c->ticket, done);
}

template <int R, int STAGES, int OCC>
static int launch_mv_tma(gcx_ctx* c, int variant, const double* z, double* q, const int* done) {
    const size_t smem = (size_t)STAGES * R * MV_THREADS * sizeof(float4);
    if (!(c->mv_cfg_mask & (1u << variant))) {        // function attributes are per device: remembered per context, not per process
        CK(cudaFuncSetAttribute(k_matvec_tma<R, STAGES, OCC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        c->mv_cfg_mask |= 1u << variant;
    }
    k_matvec_tma<R, STAGES, OCC><<<c->sm_count * OCC, TMA_THREADS, smem, c->stream>>>(c->A, c->ld, (int)c->nloc, (int)(c->ld / 4), z, q + c->row0, c->ws,
                                                                                      c->ticket, done);
    return 0;
}

static int matvec(gcx_ctx* c, int variant, const double* z, double* q, const int* done) {
    if (variant == 100) return matvec_sym(c, z, q, done);
    {
        ProfScope ps(c, 0);
        switch (variant) {
            case 0: launch_mv<8, 2>(c, z, q, done); break;
            case 1: launch_mv<8, 3>(c, z, q, done); break;
            case 2: launch_mv<4, 4>(c, z, q, done); break;
            case 3: launch_mv<4, 6>(c, z, q, done); break;
            case 4: launch_mv<16, 1>(c, z, q, done); break;
            case 5: launch_mv<2, 8>(c, z, q, done); break;
            case 6: launch_mv<8, 4>(c, z, q, done); break;
            case 7: CKR((launch_mv_tma<4, 6, 2>(c, 7, z, q, done))); break;      // 96 KB/CTA, 2 CTAs/SM
            case 8: CKR((launch_mv_tma<4, 12, 1>(c, 8, z, q, done))); break;     // 192 KB/CTA, 1 CTA/SM
            case 9: CKR((launch_mv_tma<8, 6, 1>(c, 9, z, q, done))); break;      // 192 KB/CTA, 1 CTA/SM, 8-row bands
            case 10: CKR((launch_mv_tma<4, 4, 3>(c, 10, z, q, done))); break;     // 64 KB/CTA, 3 CTAs/SM
            case 11: CKR((launch_mv_tma<2, 8, 3>(c, 11, z, q, done))); break;     // 64 KB/CTA, 3 CTAs/SM, 2-row bands
            case 12: CKR((launch_mv_tma<4, 7, 2>(c, 12, z, q, done))); break;     // 112 KB/CTA, 2 CTAs/SM
            case 13: CKR((launch_mv_tma<8, 3, 2>(c, 13, z, q, done))); break;     // 96 KB/CTA, 2 CTAs/SM, 8-row bands
            case 14: CKR((launch_mv_tma<2, 12, 2>(c, 14, z, q, done))); break;    // 96 KB/CTA, 2 CTAs/SM, 2-row bands
            case 15: CKR((launch_mv_tma<4, 5, 2>(c, 15, z, q, done))); break;     // 80 KB/CTA, 2 CTAs/SM
            case 16: CKR((launch_mv_tma<2, 14, 2>(c, 16, z, q, done))); break;    // 112 KB/CTA, 2 CTAs/SM, 2-row bands
            case 17: CKR((launch_mv_tma<1, 24, 2>(c, 17, z, q, done))); break;    // 96 KB/CTA, 2 CTAs/SM, single rows
            default: return fail("unknown matvec variant %d", variant);
        }
    }
    CK(cudaGetLastError());
    return 0;
}

extern "C" int gcx_bench_matvec(gcx_ctx* c, int variant, int iters, double* ms_total) {
    CKR(need_map(c));
    if (!ms_total || iters < 1) return fail("gcx_bench_matvec: bad arguments");
    if (variant < 0) variant = c->mv_variant;
    if (variant == 100) {
        bool ok = false;
        CKR(sym_ready(c, &ok));
        if (!ok) return fail("symmetric pass not available for this map (GCX_SYM=0, sharded, n < 2048 or asymmetric)");
    }
    // z = ones on the real columns
    std::vector<double> ones((size_t)c->nv, 0.0);
    for (int64_t i = 0; i < c->n; ++i) ones[i] = 1.0;
    CK(cudaMemcpyAsync(c->vz, ones.data(), sizeof(double) * c->nv, cudaMemcpyHostToDevice, c->stream));
    CKR(matvec(c, variant, c->vz, c->vq, nullptr));   // warm
    CK(cudaEventRecord(c->run_a, c->stream));
    for (int i = 0; i < iters; ++i) CKR(matvec(c, variant, c->vz, c->vq, nullptr));
    CK(cudaEventRecord(c->run_b, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, c->run_a, c->run_b));
    *ms_total = ms;
    return 0;
}

// ---------------------------------------------------------------------------------------------------
// iteration drivers: chunks of (matvec, allgather, vector step) are enqueued one chunk ahead of the
// host's convergence check, so the GPU never waits for the host; launches after convergence exit at once.
// ---------------------------------------------------------------------------------------------------
static int chunk_len(gcx_ctx* c) {
    // must be the same on every rank (each chunk carries collectives): derived from global quantities only
    const double est_ms = 4.0 * (double)c->n * (double)c->ld / (double)c->world / 5.0e9;   // ~5 TB/s per rank
    int ch = (int)(0.25 / std::max(est_ms, 1e-4)) + 1;
    return std::max(2, std::min(ch, 16));
}

template <typename State, typename StepFn>
static int iterate(gcx_ctx* c, State* dev_state, int max_matvecs, StepFn step, State* final_state) {
    const int CH = chunk_len(c);
    State* hs = (State*)c->h_state;   // two pinned slots
    static_assert(sizeof(State) <= 1024 && sizeof(State) % 4 == 0, "state too large / not word sized");
    State* slot[2] = {hs, (State*)((char*)c->h_state + 2048)};
    cudaEvent_t ev[2];
    CK(cudaEventCreateWithFlags(&ev[0], cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&ev[1], cudaEventDisableTiming));
    CK(cudaEventRecord(c->run_a, c->stream));
    int issued = 0, rc = 0;
    auto enqueue_chunk = [&](int b) -> int {
        for (int i = 0; i < CH; ++i) CKR(step());
        issued += CH;
        k_publish_state<<<1, 64, 0, c->stream>>>((const int*)dev_state, (volatile int*)slot[b], (int)(sizeof(State) / 4));
        CK(cudaGetLastError());
        CK(cudaEventRecord(ev[b], c->stream));
        return 0;
    };
    rc = enqueue_chunk(0);
    int cur = 0;
    while (rc == 0) {
        rc = enqueue_chunk(cur ^ 1);
        if (rc) break;
        cudaError_t e = cudaEventSynchronize(ev[cur]);
        if (e != cudaSuccess) { rc = fail("cudaEventSynchronize: %s", cudaGetErrorString(e)); break; }
        if (slot[cur]->done) break;
        if (issued > max_matvecs + 2 * CH) { rc = fail("iteration did not terminate within %d matrix passes", max_matvecs); break; }
        cur ^= 1;
    }
    cudaError_t e2 = cudaEventRecord(c->run_b, c->stream);
    k_publish_state<<<1, 64, 0, c->stream>>>((const int*)dev_state, (volatile int*)slot[0], (int)(sizeof(State) / 4));
    cudaError_t e3 = cudaStreamSynchronize(c->stream);
    cudaEventDestroy(ev[0]);
    cudaEventDestroy(ev[1]);
    if (rc) return rc;
    CK(e2);
    CK(e3);
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, c->run_a, c->run_b));
    c->last_run_ms = ms;
    memcpy(final_state, slot[0], sizeof(State));
    return 0;
}

extern "C" int gcx_ic_run(gcx_ctx* c, double tol, int iteration, double* bias, int* passes, double* last_l1, int* matvecs) {
    CKR(need_map(c));
    if (!c->have_mask) return fail("gcx_ic_run: call gcx_set_mask first");
    const int n = (int)c->n;
    bool sym = false;
    CKR(sym_ready(c, &sym));
    c->last_sym = sym;
    // the deferred-scalar pass (default on the upper-triangle path); sharded it needs the peer-mapped exchange buffers
    bool defer = sym && c->ic_fused && c->ic_defer && (c->world == 1 || c->xchg_ok);
  rerun:
    k_ic_init<<<32, EP_THREADS, 0, c->stream>>>(c->ic_state, (int)c->nv, c->keep, c->vz, c->vB, tol, iteration);
    CK(cudaGetLastError());
    c->launches += 1;
    int n4 = (int)(c->ld / 4), nloc = (int)c->nloc, row0 = (int)c->row0;
    long long nvll = c->nv;
    double* tring = c->ic_ring;
    double* uring = c->ic_ring + 3 * c->nv;
    if (defer) {
        // u_0 = the mask (k_ic_init left it in vz); the in-kernel reset of the chunk counter needs a zero to start from
        CK(cudaMemcpyAsync(uring, c->vz, sizeof(double) * c->nv, cudaMemcpyDeviceToDevice, c->stream));
        if (c->sym_ndyn > 0) CK(cudaMemsetAsync(c->sym_dyn_counter, 0, sizeof(unsigned int), c->stream));
    }
    int npass_launch = 1;
    auto step_symd = [&]() -> int {
        SymTab tab = sym_tab(c);
        IcXchg X;
        memset(&X, 0, sizeof X);
        X.world = 1;
        if (c->world > 1) X = xchg_args(c);
        void* args[] = {(void*)&c->A, (void*)&c->ld, (void*)&n, (void*)&n4, (void*)&nvll, (void*)&c->keep, (void*)&tring, (void*)&uring, (void*)&c->vB,
                        (void*)&c->rowpart, (void*)&c->colpart, (void*)&tab, (void*)&c->ticket, (void*)&c->ic_state, (void*)&c->part, (void*)&X,
                        (void*)&npass_launch};
        const size_t smem = SYM_STAGES * SYM_R * MV_THREADS * sizeof(float4);
        ProfScope ps(c, 0);
        if (c->world == 1)
            CK(cudaLaunchCooperativeKernel((void*)k_ic_pass_symd<SYM_STAGES, SYM_OCC, false>, dim3(c->sym_grid), dim3(SYM_THREADS), args, smem, c->stream));
        else
            CK(cudaLaunchCooperativeKernel((void*)k_ic_pass_symd<SYM_STAGES, SYM_OCC, true>, dim3(c->sym_grid), dim3(SYM_THREADS), args, smem, c->stream));
        return 0;
    };
    auto step_split = [&]() -> int {
        CKR(matvec(c, c->mv_variant, c->vz, c->vq, &c->ic_state->done));
        CKR(allgather_f64(c, c->vq));
        {
            ProfScope ps(c, 6);
            k_ic_epilogue<<<EP_CL, EP_THREADS, 0, c->stream>>>(c->ic_state, n, c->keep, c->vq, c->vz, c->vB);
        }
        CK(cudaGetLastError());
        return 0;
    };
    auto step_fused = [&]() -> int {
        void* args[] = {(void*)&c->A, (void*)&c->ld, (void*)&nloc, (void*)&n4, (void*)&n, (void*)&row0, (void*)&c->keep, (void*)&c->vq,
                        (void*)&c->vz, (void*)&c->vB, (void*)&c->ws, (void*)&c->ticket, (void*)&c->ic_state, (void*)&c->part};
        {
            ProfScope ps(c, 0);
            if (c->ic_fused == 2)
                CK(cudaLaunchCooperativeKernel((void*)k_ic_pass_tma<4, 6, 2>, dim3(c->ic_grid_tma), dim3(TMA_THREADS), args,
                                               6 * 4 * MV_THREADS * sizeof(float4), c->stream));
            else
                CK(cudaLaunchCooperativeKernel((void*)k_ic_pass<4, 4>, dim3(c->ic_grid), dim3(MV_THREADS), args, 0, c->stream));
        }
        CKR(allgather_f64(c, c->vq));
        return 0;
    };
    auto step_split_sym = [&]() -> int {
        CKR(matvec_sym(c, c->vz, c->vq, &c->ic_state->done));
        {
            ProfScope ps(c, 6);
            k_ic_epilogue<<<EP_CL, EP_THREADS, 0, c->stream>>>(c->ic_state, n, c->keep, c->vq, c->vz, c->vB);
        }
        CK(cudaGetLastError());
        return 0;
    };
    auto step_fused_sym = [&]() -> int {
        SymTab tab = sym_tab(c);
        void* args[] = {(void*)&c->A, (void*)&c->ld, (void*)&n, (void*)&n4, (void*)&c->keep, (void*)&c->vq, (void*)&c->vz, (void*)&c->vB,
                        (void*)&c->rowpart, (void*)&c->colpart, (void*)&tab, (void*)&c->ticket, (void*)&c->ic_state, (void*)&c->part};
        const size_t smem = SYM_STAGES * SYM_R * MV_THREADS * sizeof(float4);
        if (c->sym_ndyn > 0) CK(cudaMemsetAsync(c->sym_dyn_counter, 0, sizeof(unsigned int), c->stream));
        {
            ProfScope ps(c, 0);
            if (c->world == 1) {
                CK(cudaLaunchCooperativeKernel((void*)k_ic_pass_sym_tma<SYM_STAGES, SYM_OCC, true>, dim3(c->sym_grid), dim3(SYM_THREADS), args, smem, c->stream));
            } else {
                // sharded: the fold runs as its own kernel so that the ranks' partial t can be added before the next launch
                CK(cudaLaunchCooperativeKernel((void*)k_ic_pass_sym_tma<SYM_STAGES, SYM_OCC, false>, dim3(c->sym_grid), dim3(SYM_THREADS), args, smem, c->stream));
                k_sym_fold<<<(n * 8 + 255) / 256, 256, 0, c->stream>>>(n, n4, c->rowpart, c->colpart, tab, c->vq, &c->ic_state->done);
            }
        }
        if (c->world > 1) {
            CK(cudaGetLastError());
            c->launches += 1;
            CKR(allreduce_f64(c, c->vq, (size_t)c->n));
        }
        return 0;
    };
    auto step = [&]() -> int {
        if (defer) return step_symd();
        if (sym) return c->ic_fused ? step_fused_sym() : step_split_sym();
        return c->ic_fused ? step_fused() : step_split();
    };
    IcState fin;
    int rc_it = 0;
    if (defer && c->ic_persist) {
        // the whole solve is one cooperative launch: passes are separated by grid barriers, convergence is decided on the device
        npass_launch = iteration + 8;
        rc_it = [&]() -> int {
            CK(cudaEventRecord(c->run_a, c->stream));
            CKR(step_symd());
            CK(cudaEventRecord(c->run_b, c->stream));
            CKR(read_small(c, c->ic_state, sizeof(IcState), &fin));
            float ms = 0;
            CK(cudaEventElapsedTime(&ms, c->run_a, c->run_b));
            c->last_run_ms = ms;
            if (!fin.done) return fail("gcx_ic_run: the persistent pass ended without a verdict");
            return 0;
        }();
    } else {
        rc_it = iterate<IcState>(c, c->ic_state, iteration + 6, step, &fin);
    }
    if (defer && c->world > 1) c->xchg_seq += (unsigned int)iteration + 4096u;      // the same on every rank
    CKR(rc_it);
    if (defer && fin.error) return fail("gcx_ic_run: a peer rank did not deliver its partial vector within 20 s (peer exchange)");
    if (defer && fin.irregular) {
        // a bin's product changed between zero and non-zero from one pass to the next (cannot happen on a symmetric non-negative
        // map): the deferred recurrence does not cover it -- rerun in the literal order of lib/normalizeCore.pyx:42-53
        defer = false;
        goto rerun;
    }
    c->last_matvecs_issued = fin.matvecs;
    k_ic_scale<<<(int)((c->nv + 255) / 256), 256, 0, c->stream>>>((int)c->nv, n, c->keep, c->vB, c->vscale);
    CK(cudaGetLastError());
    c->launches += 1;
    c->have_scale = true;
    if (bias) CK(cudaMemcpyAsync(bias, c->vB, sizeof(double) * n, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    if (passes) *passes = fin.pass;
    if (last_l1) *last_l1 = fin.l1;
    if (matvecs) *matvecs = fin.pass + 1;          // products the recurrence needs; gcx_last_run_products gives the launched count
    return 0;
}

extern "C" int gcx_kr_run(gcx_ctx* c, double tol, double delta, double Delta, double* x, int* outer, int* mvp, double* rout) {
    CKR(need_map(c));
    if (!c->have_mask) return fail("gcx_kr_run: call gcx_set_mask first");
    const int n = (int)c->n;
    k_kr_init<<<32, EP_THREADS, 0, c->stream>>>(c->kr_state, (int)c->nv, c->keep, c->kr, tol, delta, Delta);
    CK(cudaGetLastError());
    c->launches += 1;
    bool sym = false;
    CKR(sym_ready(c, &sym));
    c->last_sym = sym;
    auto step = [&]() -> int {
        CKR(matvec(c, sym ? 100 : c->mv_variant, c->vz, c->vq, &c->kr_state->done));
        if (!sym) CKR(allgather_f64(c, c->vq));      // the symmetric pass already left the complete vector (allreduce)
        {
            ProfScope ps(c, 6);
            k_kr_step<<<EP_CL, EP_THREADS, 0, c->stream>>>(c->kr_state, n, c->keep, c->vq, c->kr);
        }
        CK(cudaGetLastError());
        return 0;
    };
    KrState fin;
    CKR(iterate<KrState>(c, c->kr_state, 20000, step, &fin));
    c->last_matvecs_issued = fin.matvecs;
    CK(cudaMemcpyAsync(c->vscale, c->kr.x, sizeof(double) * c->nv, cudaMemcpyDeviceToDevice, c->stream));
    c->have_scale = true;
    if (x) CK(cudaMemcpyAsync(x, c->kr.x, sizeof(double) * n, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    if (outer) *outer = fin.outer;
    if (mvp) *mvp = fin.mvp;
    if (rout) *rout = fin.rout;
    return 0;
}

// ---------------------------------------------------------------------------------------------------
// elementwise passes
// ---------------------------------------------------------------------------------------------------
static int mm_begin(gcx_ctx* c) {
    MinMax init{0xffffffffu, 0u};
    memcpy(c->h_state, &init, sizeof init);
    CK(cudaMemcpyAsync(c->mm, c->h_state, sizeof init, cudaMemcpyHostToDevice, c->stream));
    return 0;
}

static int mm_end(gcx_ctx* c, float* minv, float* maxv) {
    if (c->world > 1) {
        NK(g_nccl.AllReduce(&c->mm->min_nz, &c->mm->min_nz, 1, ncclUint32, ncclMin, c->comm, c->stream));
        NK(g_nccl.AllReduce(&c->mm->max_all, &c->mm->max_all, 1, ncclUint32, ncclMax, c->comm, c->stream));
    }
    MinMax r;
    CKR(read_small(c, c->mm, sizeof r, &r));
    const float nanv = __builtin_nanf("");
    if (minv) *minv = r.min_nz == 0xffffffffu ? nanv : ord2f(r.min_nz);
    if (maxv) *maxv = r.max_all == 0u ? nanv : ord2f(r.max_all);
    return 0;
}

// elementwise kernels: (rows per band, CTAs per SM) variants, selectable with GCX_EW_VARIANT for sweeps
static int ew_variant() {
    static int v = -1;
    if (v < 0) { const char* e = getenv("GCX_EW_VARIANT"); v = e ? atoi(e) : 3; }   // default R=4, 4 CTAs/SM (profiles/r01_ew_variants.txt)
    return v;
}
#define EW_DISPATCH(KERNEL, ...)                                                                                   \
    switch (ew_variant()) {                                                                                        \
        case 1: KERNEL<8, 2><<<c->sm_count * 2, MV_THREADS, 0, c->stream>>>(__VA_ARGS__); break;                    \
        case 2: KERNEL<8, 3><<<c->sm_count * 3, MV_THREADS, 0, c->stream>>>(__VA_ARGS__); break;                    \
        case 3: KERNEL<4, 4><<<c->sm_count * 4, MV_THREADS, 0, c->stream>>>(__VA_ARGS__); break;                    \
        case 4: KERNEL<4, 6><<<c->sm_count * 6, MV_THREADS, 0, c->stream>>>(__VA_ARGS__); break;                    \
        case 5: KERNEL<2, 8><<<c->sm_count * 8, MV_THREADS, 0, c->stream>>>(__VA_ARGS__); break;                    \
        default: KERNEL<4, 3><<<c->sm_count * 3, MV_THREADS, 0, c->stream>>>(__VA_ARGS__); break;                   \
    }

extern "C" int gcx_apply_sym_scale(gcx_ctx* c, const double* s, int masked_mode, int in_place, float* minv, float* maxv) {
    CKR(need_map(c));
    CKR(fence_copy(c));
    if (!c->have_mask) return fail("gcx_apply_sym_scale: call gcx_set_mask first");
    if (s) {
        CK(cudaMemsetAsync(c->vscale, 0, sizeof(double) * c->nv, c->stream));
        CK(cudaMemcpyAsync(c->vscale, s, sizeof(double) * c->n, cudaMemcpyHostToDevice, c->stream));
        c->have_scale = true;
    }
    if (!c->have_scale) return fail("gcx_apply_sym_scale: no scale vector (pass s or run gcx_ic_run / gcx_kr_run)");
    float* out = c->A;
    if (!in_place) { CKR(ensure_out(c)); out = c->Out; }
    CKR(mm_begin(c));
    {
        ProfScope ps(c, 1);
            EW_DISPATCH(k_apply_sym, c->A, out, c->ld, (int)c->nloc, (int)(c->ld / 4), (int)c->row0, c->vscale, c->keep, masked_mode, c->mm)
    }
    CK(cudaGetLastError());
    if (in_place) { c->have_rowsum = false; c->sym_state = -1; }
    return mm_end(c, minv, maxv);
}

static int mm_read(gcx_ctx* c, MinMax* dev, float* minv, float* maxv) {
    if (c->world > 1) {
        NK(g_nccl.AllReduce(&dev->min_nz, &dev->min_nz, 1, ncclUint32, ncclMin, c->comm, c->stream));
        NK(g_nccl.AllReduce(&dev->max_all, &dev->max_all, 1, ncclUint32, ncclMax, c->comm, c->stream));
    }
    MinMax r;
    CKR(read_small(c, dev, sizeof r, &r));
    const float nanv = __builtin_nanf("");
    if (minv) *minv = r.min_nz == 0xffffffffu ? nanv : ord2f(r.min_nz);
    if (maxv) *maxv = r.max_all == 0u ? nanv : ord2f(r.max_all);
    return 0;
}

// .hic norm-vector application at import (lib/hic2gcmap.py:82-98): c1 indexes rows (bin_x), c2 columns (bin_y); n doubles each
extern "C" int gcx_apply_norm_vectors(gcx_ctx* c, const double* c1, const double* c2, int in_place, float* minv, float* maxv) {
    CKR(need_map(c));
    CKR(fence_copy(c));
    if (!c1 || !c2) return fail("gcx_apply_norm_vectors: null norm vector");
    // the two vectors borrow the KR work vectors (any resident scale vector is unaffected)
    double* d1 = c->kr.w;
    double* d2 = c->kr.p;
    CK(cudaMemsetAsync(d1, 0, sizeof(double) * c->nv, c->stream));
    CK(cudaMemsetAsync(d2, 0, sizeof(double) * c->nv, c->stream));
    CK(cudaMemcpyAsync(d1, c1, sizeof(double) * c->n, cudaMemcpyHostToDevice, c->stream));
    CK(cudaMemcpyAsync(d2, c2, sizeof(double) * c->n, cudaMemcpyHostToDevice, c->stream));
    float* out = c->A;
    if (!in_place) { CKR(ensure_out(c)); out = c->Out; }
    CKR(mm_begin(c));
    {
        ProfScope ps(c, 1);
        EW_DISPATCH(k_apply_norm_vectors, c->A, out, c->ld, (int)c->nloc, (int)(c->ld / 4), (int)c->n, (int)c->row0, d1, d2, c->mm)
    }
    CK(cudaGetLastError());
    if (in_place) { c->have_rowsum = false; c->sym_state = -1; }
    return mm_end(c, minv, maxv);
}

__global__ void k_f64_to_f32(int nv, const double* __restrict__ a, float* __restrict__ b);

// Fused tail of a pipeline that wants BOTH normalizations of one map (bench.py's step; the batch driver): the scaled map of
// gcx_apply_sym_scale (resident scale vector and mask) into output 1 and the vanilla-coverage map of gcx_vc_run (mask keep_vc)
// into output 2, from ONE read of the input: 12 n^2 bytes of HBM traffic instead of 16 n^2.
extern "C" int gcx_apply_sym_vc(gcx_ctx* c, const uint8_t* keep_vc, int masked_mode, int sqroot, float* min_s, float* max_s, float* min_v,
                                float* max_v) {
    CKR(need_map(c));
    CKR(fence_copy(c));
    if (!c->have_mask || !c->have_scale) return fail("gcx_apply_sym_vc: needs the mask and the scale vector of a finished gcx_ic_run / gcx_kr_run");
    if (!keep_vc) return fail("gcx_apply_sym_vc: null VC mask");
    CKR(run_row_stats(c));
    k_f64_to_f32<<<(int)((c->nv + 255) / 256), 256, 0, c->stream>>>((int)c->nv, c->vrowsum, c->csum32);
    CK(cudaGetLastError());
    c->launches += 1;
    CKR(ensure_out(c));
    if (!c->Out2) CK(cudaMalloc(&c->Out2, sizeof(float) * (size_t)c->nloc * c->ld));
    if (!c->keep2) CK(cudaMalloc(&c->keep2, (size_t)c->nv));
    CK(cudaMemsetAsync(c->keep2, 0, (size_t)c->nv, c->stream));
    CK(cudaMemcpyAsync(c->keep2, keep_vc, (size_t)c->n, cudaMemcpyHostToDevice, c->stream));
    MinMax init{0xffffffffu, 0u};
    memcpy(c->h_state, &init, sizeof init);
    CK(cudaMemcpyAsync(c->mm, c->h_state, sizeof init, cudaMemcpyHostToDevice, c->stream));
    CK(cudaMemcpyAsync(c->mm2, c->h_state, sizeof init, cudaMemcpyHostToDevice, c->stream));
    {
        // two results per element need more registers than the single-result passes: 3 CTAs/SM without spills is the default
        static const int v2 = [] { const char* e = getenv("GCX_EW2_VARIANT"); return e ? atoi(e) : 1; }();
        ProfScope ps(c, 1);
#define EW2(RR, OC) k_apply_sym_vc<RR, OC><<<c->sm_count * OC, MV_THREADS, 0, c->stream>>>(c->A, c->Out, c->Out2, c->ld, (int)c->nloc, (int)(c->ld / 4), (int)c->n, \
        (int)c->row0, c->vscale, c->keep, masked_mode, c->csum32, c->keep2, sqroot, c->mm, c->mm2)
        switch (v2) {
            case 1: EW2(4, 4); break;
            case 2: EW2(2, 6); break;
            case 3: EW2(8, 2); break;
            case 4: EW2(2, 4); break;
            default: EW2(4, 3); break;
        }
#undef EW2
    }
    CK(cudaGetLastError());
    CKR(mm_read(c, c->mm, min_s, max_s));
    return mm_read(c, c->mm2, min_v, max_v);
}

__global__ void k_f64_to_f32(int nv, const double* __restrict__ a, float* __restrict__ b) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nv; i += gridDim.x * blockDim.x) b[i] = (float)a[i];
}

extern "C" int gcx_vc_run(gcx_ctx* c, int sqroot, int in_place, float* minv, float* maxv, float* csum) {
    CKR(need_map(c));
    CKR(fence_copy(c));
    if (!c->have_mask) return fail("gcx_vc_run: call gcx_set_mask first");
    CKR(run_row_stats(c));
    // csum = np.sum(A, axis=1) is float32 in the reference (lib/normalizeCore.pyx:72)
    k_f64_to_f32<<<(int)((c->nv + 255) / 256), 256, 0, c->stream>>>((int)c->nv, c->vrowsum, c->csum32);
    CK(cudaGetLastError());
    c->launches += 1;
    float* out = c->A;
    if (!in_place) { CKR(ensure_out(c)); out = c->Out; }
    CKR(mm_begin(c));
    {
        ProfScope ps(c, 2);
        EW_DISPATCH(k_vc_apply, c->A, out, c->ld, (int)c->nloc, (int)(c->ld / 4), (int)c->n, (int)c->row0, c->csum32, c->keep, sqroot, c->mm)
    }
    CK(cudaGetLastError());
    if (csum) CK(cudaMemcpyAsync(csum, c->csum32, sizeof(float) * c->n, cudaMemcpyDeviceToHost, c->stream));
    if (in_place) { c->have_rowsum = false; c->sym_state = -1; }
    return mm_end(c, minv, maxv);
}

static dim3 diag_grid(gcx_ctx* c) {
    return dim3((unsigned)((c->n + MV_THREADS - 1) / MV_THREADS), (unsigned)((c->nloc + DIAG_RB - 1) / DIAG_RB));
}

static dim3 diag_hist_grid(gcx_ctx* c) {
    return dim3((unsigned)((c->n + MV_THREADS - 1) / MV_THREADS), (unsigned)((c->nloc + DIAG_HB - 1) / DIAG_HB));
}

static size_t align256(size_t x) { return (x + 255) / 256 * 256; }

static int diag_mean(gcx_ctx* c) {
    const int n = (int)c->n, nbands = (int)((c->nloc + DIAG_RB - 1) / DIAG_RB);
    const size_t b_psum = align256(sizeof(double) * (size_t)nbands * n), b_pcnt = align256(sizeof(int) * (size_t)nbands * n),
                 b_sc = align256(sizeof(double) * 2 * (size_t)c->nv);
    CKR(ensure_scratch(c, b_psum + b_pcnt + b_sc));
    double* psum = (double*)c->scratch;
    int* pcnt = (int*)((char*)c->scratch + b_psum);
    double* sumcnt = (double*)((char*)c->scratch + b_psum + b_pcnt);     // [2][nv]
    {
        ProfScope ps(c, 4);
        k_diag_mean_partial<<<diag_grid(c), MV_THREADS, 0, c->stream>>>(c->A, c->ld, (int)c->nloc, n, (int)c->row0, psum, pcnt);
    }
    CK(cudaGetLastError());
    k_diag_mean_reduce<<<(n + 63) / 64, 64 * DIAG_RED_SL, 0, c->stream>>>(nbands, n, (int)c->row0, psum, pcnt, sumcnt, sumcnt + c->nv);
    CK(cudaGetLastError());
    c->launches += 1;
    if (c->world > 1) NK(g_nccl.AllReduce(sumcnt, sumcnt, 2 * (size_t)c->nv, ncclFloat64, ncclSum, c->comm, c->stream));
    k_diag_mean_final<<<(n + 255) / 256, 256, 0, c->stream>>>(n, sumcnt, sumcnt + c->nv, c->vavg);
    CK(cudaGetLastError());
    c->launches += 1;
    return 0;
}

static int diag_median(gcx_ctx* c) {
    const int n = (int)c->n;
    const size_t b_hist = align256(sizeof(unsigned int) * DIAG_NBIN * (size_t)n),
                 b_small = align256(sizeof(unsigned int) * (DIAG_SEL_ARRAYS * (size_t)n + 1));
    CKR(ensure_scratch(c, b_hist + b_small));
    DiagSel S{};
    unsigned int* small = (unsigned int*)((char*)c->scratch + b_hist);     // prefix, krem, cnt, need, next, nbp, dirty, done | pending
    diag_sel_bind(S, (unsigned int*)c->scratch, small, (size_t)n, small + DIAG_SEL_ARRAYS * (size_t)n);
    CK(cudaMemsetAsync(c->scratch, 0, b_hist + b_small, c->stream));
    const int sel_grid = (int)(((long long)n * 32 + MV_THREADS - 1) / MV_THREADS);
    for (int pass = 0; pass < 4; ++pass) {
        {
            ProfScope ps(c, 4);
            const dim3 g = diag_hist_grid(c);
            if (pass == 0) k_diag_hist<0><<<g, MV_THREADS, 0, c->stream>>>(c->A, c->ld, (int)c->nloc, n, (int)c->row0, S);
            else if (pass == 1) k_diag_hist<1><<<g, MV_THREADS, 0, c->stream>>>(c->A, c->ld, (int)c->nloc, n, (int)c->row0, S);
            else if (pass == 2) k_diag_hist<2><<<g, MV_THREADS, 0, c->stream>>>(c->A, c->ld, (int)c->nloc, n, (int)c->row0, S);
            else k_diag_hist<3><<<g, MV_THREADS, 0, c->stream>>>(c->A, c->ld, (int)c->nloc, n, (int)c->row0, S);
        }
        CK(cudaGetLastError());
        if (c->world > 1) {
            // sharded: the histograms add up over the ranks; every rank then makes the same selection
            NK(g_nccl.AllReduce(S.hist, S.hist, DIAG_NBIN * (size_t)n, ncclUint32, ncclSum, c->comm, c->stream));
            if (pass == 1) NK(g_nccl.AllReduce(S.dirty, S.dirty, (size_t)n, ncclUint32, ncclMax, c->comm, c->stream));
        }
        k_diag_select<<<sel_grid, MV_THREADS, 0, c->stream>>>(n, pass, S);
        CK(cudaGetLastError());
        c->launches += 1;
    }
    if (c->world > 1) NK(g_nccl.AllReduce(S.next, S.next, (size_t)n, ncclUint32, ncclMin, c->comm, c->stream));
    k_diag_median_final<<<(n + 255) / 256, 256, 0, c->stream>>>(n, S, c->vavg);
    CK(cudaGetLastError());
    c->launches += 1;
    return 0;
}

extern "C" int gcx_diag_stats(gcx_ctx* c, int median, double* avg) {
    CKR(need_map(c));
    CK(cudaMemsetAsync(c->vavg, 0, sizeof(double) * c->nv, c->stream));
    CKR(median ? diag_median(c) : diag_mean(c));
    if (avg) CK(cudaMemcpyAsync(avg, c->vavg, sizeof(double) * c->n, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    return 0;
}

extern "C" int gcx_diag_apply(gcx_ctx* c, const double* avg, int subtract, int in_place, float* minv, float* maxv) {
    CKR(need_map(c));
    CKR(fence_copy(c));
    if (avg) {
        CK(cudaMemsetAsync(c->vavg, 0, sizeof(double) * c->nv, c->stream));
        CK(cudaMemcpyAsync(c->vavg, avg, sizeof(double) * c->n, cudaMemcpyHostToDevice, c->stream));
    }
    float* out = c->A;
    if (!in_place) { CKR(ensure_out(c)); out = c->Out; }
    CKR(mm_begin(c));
    {
        ProfScope ps(c, 5);
        k_diag_recip<<<(int)((c->n + 255) / 256), 256, 0, c->stream>>>((int)c->n, c->vavg, c->vavr);
        static const int lane = getenv("GCX_DIAG_APPLY_LANE") ? atoi(getenv("GCX_DIAG_APPLY_LANE")) : 1;
        if (!lane) {
            EW_DISPATCH(k_diag_apply, c->A, out, c->ld, (int)c->nloc, (int)(c->ld / 4), (int)c->n, (int)c->row0, c->vavr, subtract, c->mm)
        } else if (subtract) {
            k_diag_apply_lane<4, 4, true><<<c->sm_count * 4, MV_THREADS, 0, c->stream>>>(c->A, out, c->ld, (int)c->nloc, (int)c->n, (int)c->row0, c->vavr, c->mm);
        } else {
            k_diag_apply_lane<4, 4, false><<<c->sm_count * 4, MV_THREADS, 0, c->stream>>>(c->A, out, c->ld, (int)c->nloc, (int)c->n, (int)c->row0, c->vavr, c->mm);
        }
    }
    CK(cudaGetLastError());
    if (in_place) { c->have_rowsum = false; c->sym_state = -1; }
    return mm_end(c, minv, maxv);
}

// ---------------------------------------------------------------------------------------------------
// batched distance-frequency normalization over many resident maps (one grid per stage for the whole set)
// ---------------------------------------------------------------------------------------------------
static void launch_diag_hist_b(int pass, unsigned grid, cudaStream_t st, const BMap* dt, int count, int nblocks) {
    if (pass == 0) k_diag_hist_b<0><<<grid, MV_THREADS, 0, st>>>(dt, count, nblocks);
    else if (pass == 1) k_diag_hist_b<1><<<grid, MV_THREADS, 0, st>>>(dt, count, nblocks);
    else if (pass == 2) k_diag_hist_b<2><<<grid, MV_THREADS, 0, st>>>(dt, count, nblocks);
    else k_diag_hist_b<3><<<grid, MV_THREADS, 0, st>>>(dt, count, nblocks);
}

extern "C" int gcx_mcfs_batch(gcx_ctx** ctxs, int count, int median, int subtract, float* minv, float* maxv) {
    if (!ctxs || count < 1) return fail("gcx_mcfs_batch: no maps");
    if (count > 120) return fail("gcx_mcfs_batch: at most 120 maps per call");
    gcx_ctx* c0 = ctxs[0];
    for (int m = 0; m < count; ++m) {
        gcx_ctx* c = ctxs[m];
        if (!c || !c->A) return fail("gcx_mcfs_batch: map %d has no resident matrix", m);
        if (c->device != c0->device) return fail("gcx_mcfs_batch: all maps must be resident on one device");
        if (c->world != 1 || c->row0 != 0 || c->nloc != c->n) return fail("gcx_mcfs_batch: sharded maps are not batched");
    }
    CKR(set_dev(c0));
    // workspace of the whole batch in the first context's arena: table | per map { histograms, selection state | mean partials }
    std::vector<BMap> tab((size_t)count);
    size_t off = align256(sizeof(BMap) * (size_t)count), zero_from = off;
    std::vector<size_t> o_hist((size_t)count), o_small((size_t)count), o_psum((size_t)count), o_pcnt((size_t)count), o_sc((size_t)count);
    long long blk = 0, sel = 0, red = 0, fin = 0, units = 0;
    int apply_variant = 3;                   // rows per unit / CTAs per SM of the batched apply: 0 = 4 / 4, 1 = 8 / 3, 2 = 8 / 2
    if (const char* e = getenv("GCX_DIAG_APPLY_VARIANT")) apply_variant = atoi(e);
    const int R = (apply_variant == 0 || apply_variant == 3) ? 4 : 8;       // 3, 4: lane-strided columns, 4 / 4 and 8 / 3
    for (int m = 0; m < count; ++m) {
        gcx_ctx* c = ctxs[m];
        const int n = (int)c->n;
        BMap& M = tab[m];
        memset(&M, 0, sizeof M);
        M.ld = c->ld; M.n = n;
        M.gx = (n + MV_THREADS - 1) / MV_THREADS;
        M.gy = median ? (n + DIAG_HB - 1) / DIAG_HB : (n + DIAG_RB - 1) / DIAG_RB;      // row bands of the histogram / the mean-partial grid
        M.blk0 = (int)blk; M.sel0 = (int)sel; M.red0 = (int)red; M.fin0 = (int)fin; M.u0 = units;
        blk += (long long)M.gx * M.gy;
        sel += ((long long)n * 32 + MV_THREADS - 1) / MV_THREADS;
        red += (n + 63) / 64;
        fin += (n + 255) / 256;
        units += (long long)((n + R - 1) / R) * (((int)(c->ld / 4) + MV_THREADS - 1) / MV_THREADS);
        if (median) {
            o_hist[m] = off; off += align256(sizeof(unsigned int) * DIAG_NBIN * (size_t)n);
            o_small[m] = off; off += align256(sizeof(unsigned int) * DIAG_SEL_ARRAYS * (size_t)n);
        } else {
            o_psum[m] = off; off += align256(sizeof(double) * (size_t)M.gy * n);
            o_pcnt[m] = off; off += align256(sizeof(int) * (size_t)M.gy * n);
            o_sc[m] = off; off += align256(sizeof(double) * 2 * (size_t)n);
        }
    }
    if (blk > 0x7fffffffLL || sel > 0x7fffffffLL) return fail("gcx_mcfs_batch: batch too large");
    const size_t o_pending = off;             // one word for the whole batch: diagonals that need the two narrow passes
    off += 256;
    const size_t o_mm = off;
    off += align256(sizeof(MinMax) * (size_t)count);
    // the batch runs on the first context's stream: it waits for whatever the other contexts have in flight ...
    for (int m = 1; m < count; ++m) {
        CK(cudaEventRecord(ctxs[m]->ev_ready, ctxs[m]->stream));
        CK(cudaStreamWaitEvent(c0->stream, ctxs[m]->ev_ready, 0));
    }
    CKR(ensure_scratch(c0, off));
    char* base = (char*)c0->scratch;
    for (int m = 0; m < count; ++m) {
        gcx_ctx* c = ctxs[m];
        CKR(fence_copy(c));
        CKR(ensure_out(c));
        BMap& M = tab[m];
        M.A = c->A; M.Out = c->Out; M.avg = c->vavg; M.avr = c->vavr; M.mm = c->mm;
        if (median) {
            const size_t n = (size_t)c->n;
            diag_sel_bind(M.S, (unsigned int*)(base + o_hist[m]), (unsigned int*)(base + o_small[m]), n, (unsigned int*)(base + o_pending));
        } else {
            M.psum = (double*)(base + o_psum[m]);
            M.pcnt = (int*)(base + o_pcnt[m]);
            M.sum = (double*)(base + o_sc[m]);
            M.cnt = M.sum + c->n;
        }
    }
    cudaStream_t st = c0->stream;
    // the table travels through the (pageable) vector: synchronous with respect to the host, ordered on the stream
    CK(cudaMemcpyAsync(base, tab.data(), sizeof(BMap) * (size_t)count, cudaMemcpyHostToDevice, st));
    CK(cudaStreamSynchronize(st));
    const BMap* dt = (const BMap*)base;
    if (median) CK(cudaMemsetAsync(base + zero_from, 0, o_mm - zero_from, st));
    k_mm_init_b<<<1, 128, 0, st>>>(dt, count);
    if (median) {
        const long long cap = (long long)c0->sm_count * 16;
        for (int pass = 0; pass < 4; ++pass) {
            const unsigned gh = (unsigned)(pass < 2 ? blk : std::min(blk, cap)), gs = (unsigned)(pass < 2 ? sel : std::min(sel, cap));
            {
                ProfScope ps(c0, 4);
                launch_diag_hist_b(pass, gh, st, dt, count, (int)blk);
            }
            k_diag_select_b<<<gs, MV_THREADS, 0, st>>>(dt, count, pass, (int)sel);
            c0->launches += 1;
        }
        k_diag_median_final_b<<<(unsigned)fin, 256, 0, st>>>(dt, count);
        c0->launches += 1;
    } else {
        {
            ProfScope ps(c0, 4);
            k_diag_mean_partial_b<<<(unsigned)blk, MV_THREADS, 0, st>>>(dt, count);
        }
        k_diag_mean_reduce_b<<<(unsigned)red, 64 * DIAG_RED_SL, 0, st>>>(dt, count);
        k_diag_mean_final_b<<<(unsigned)fin, 256, 0, st>>>(dt, count);
        c0->launches += 2;
    }
    CK(cudaGetLastError());
    k_diag_recip_b<<<(unsigned)fin, 256, 0, st>>>(dt, count);
    {
        ProfScope ps(c0, 5);
        if (apply_variant == 0) k_diag_apply_b<4, 4><<<c0->sm_count * 4, MV_THREADS, 0, st>>>(dt, count, units, subtract);
        else if (apply_variant == 1) k_diag_apply_b<8, 3><<<c0->sm_count * 3, MV_THREADS, 0, st>>>(dt, count, units, subtract);
        else if (apply_variant == 2) k_diag_apply_b<8, 2><<<c0->sm_count * 2, MV_THREADS, 0, st>>>(dt, count, units, subtract);
        else if (apply_variant == 3) {
            if (subtract) k_diag_apply_lane_b<4, 4, true><<<c0->sm_count * 4, MV_THREADS, 0, st>>>(dt, count, units);
            else k_diag_apply_lane_b<4, 4, false><<<c0->sm_count * 4, MV_THREADS, 0, st>>>(dt, count, units);
        } else {
            if (subtract) k_diag_apply_lane_b<8, 3, true><<<c0->sm_count * 3, MV_THREADS, 0, st>>>(dt, count, units);
            else k_diag_apply_lane_b<8, 3, false><<<c0->sm_count * 3, MV_THREADS, 0, st>>>(dt, count, units);
        }
    }
    CK(cudaGetLastError());
    MinMax* gathered = (MinMax*)(base + o_mm);
    k_mm_gather_b<<<1, 128, 0, st>>>(dt, count, gathered);
    CK(cudaGetLastError());
    c0->launches += 2;
    // ... and whatever the other contexts enqueue next waits for the batch
    CK(cudaEventRecord(c0->ev_ready, st));
    for (int m = 1; m < count; ++m) CK(cudaStreamWaitEvent(ctxs[m]->stream, c0->ev_ready, 0));
    std::vector<MinMax> mmh((size_t)count);
    for (int m0 = 0; m0 < count; m0 += 120) {
        const int k = std::min(120, count - m0);
        CKR(read_small(c0, gathered + m0, sizeof(MinMax) * (size_t)k, mmh.data() + m0));
    }
    const float nanv = __builtin_nanf("");
    for (int m = 0; m < count; ++m) {
        if (minv) minv[m] = mmh[m].min_nz == 0xffffffffu ? nanv : ord2f(mmh[m].min_nz);
        if (maxv) maxv[m] = mmh[m].max_all == 0u ? nanv : ord2f(mmh[m].max_all);
    }
    return 0;
}

// getAvgContactByDistance for a LIST of maps (lib/cmstats.py:572-578): per distance d the non-zero entries of the d-th diagonal of
// every map that has one are pooled, then median or mean.  Median: all maps feed ONE set of per-diagonal histograms (radix select
// as for a single map); mean: per-map sums and counts, added per diagonal.
extern "C" int gcx_diag_stats_pooled(gcx_ctx** ctxs, int count, int median, double* avg, int64_t nmax) {
    if (!ctxs || count < 1 || !avg) return fail("gcx_diag_stats_pooled: bad arguments");
    if (count > 120) return fail("gcx_diag_stats_pooled: at most 120 maps per call");
    gcx_ctx* c0 = ctxs[0];
    int64_t nm = 0;
    for (int m = 0; m < count; ++m) {
        gcx_ctx* c = ctxs[m];
        if (!c || !c->A) return fail("gcx_diag_stats_pooled: map %d has no resident matrix", m);
        if (c->device != c0->device) return fail("gcx_diag_stats_pooled: all maps must be resident on one device");
        if (c->world != 1 || c->row0 != 0 || c->nloc != c->n) return fail("gcx_diag_stats_pooled: sharded maps are not pooled");
        nm = std::max(nm, c->n);
    }
    if (nmax != nm) return fail("gcx_diag_stats_pooled: the output needs %lld entries (the largest map)", (long long)nm);
    CKR(set_dev(c0));
    std::vector<BMap> tab((size_t)count + 1);
    size_t off = align256(sizeof(BMap) * ((size_t)count + 1));
    const size_t o_hist = off; off += median ? align256(sizeof(unsigned int) * DIAG_NBIN * (size_t)nm) : 0;
    const size_t o_small = off; off += median ? align256(sizeof(unsigned int) * (DIAG_SEL_ARRAYS * (size_t)nm + 1)) : 0;
    const size_t o_avg = off; off += align256(sizeof(double) * (size_t)nm);
    std::vector<size_t> o_psum((size_t)count), o_pcnt((size_t)count), o_sc((size_t)count);
    long long blk = 0, red = 0;
    for (int m = 0; m < count; ++m) {
        gcx_ctx* c = ctxs[m];
        const int n = (int)c->n;
        BMap& M = tab[m];
        memset(&M, 0, sizeof M);
        M.ld = c->ld; M.n = n;
        M.gx = (n + MV_THREADS - 1) / MV_THREADS;
        M.gy = median ? (n + DIAG_HB - 1) / DIAG_HB : (n + DIAG_RB - 1) / DIAG_RB;      // row bands of the histogram / the mean-partial grid
        M.blk0 = (int)blk; M.red0 = (int)red;
        blk += (long long)M.gx * M.gy;
        red += (n + 63) / 64;
        if (!median) {
            o_psum[m] = off; off += align256(sizeof(double) * (size_t)M.gy * n);
            o_pcnt[m] = off; off += align256(sizeof(int) * (size_t)M.gy * n);
            o_sc[m] = off; off += align256(sizeof(double) * 2 * (size_t)n);
        }
    }
    if (blk > 0x7fffffffLL) return fail("gcx_diag_stats_pooled: batch too large");
    for (int m = 1; m < count; ++m) {
        CK(cudaEventRecord(ctxs[m]->ev_ready, ctxs[m]->stream));
        CK(cudaStreamWaitEvent(c0->stream, ctxs[m]->ev_ready, 0));
    }
    CKR(ensure_scratch(c0, off));
    char* base = (char*)c0->scratch;
    DiagSel S{};
    if (median) {
        unsigned int* small = (unsigned int*)(base + o_small);
        diag_sel_bind(S, (unsigned int*)(base + o_hist), small, (size_t)nm, small + DIAG_SEL_ARRAYS * (size_t)nm);
    }
    double* davg = (double*)(base + o_avg);
    for (int m = 0; m < count; ++m) {
        gcx_ctx* c = ctxs[m];
        BMap& M = tab[m];
        M.A = c->A; M.S = S;
        if (!median) {
            M.psum = (double*)(base + o_psum[m]);
            M.pcnt = (int*)(base + o_pcnt[m]);
            M.sum = (double*)(base + o_sc[m]);
            M.cnt = M.sum + c->n;
        }
    }
    // one pseudo-map of nmax diagonals for the stages that run once for the pool (selection, final)
    BMap& P = tab[(size_t)count];
    memset(&P, 0, sizeof P);
    P.n = (int)nm; P.S = S; P.avg = davg;
    cudaStream_t st = c0->stream;
    CK(cudaMemcpyAsync(base, tab.data(), sizeof(BMap) * ((size_t)count + 1), cudaMemcpyHostToDevice, st));
    CK(cudaStreamSynchronize(st));
    const BMap* dt = (const BMap*)base;
    const BMap* dp = dt + count;
    if (median) {
        CK(cudaMemsetAsync(base + o_hist, 0, o_avg - o_hist, st));
        const unsigned sel = (unsigned)(((long long)nm * 32 + MV_THREADS - 1) / MV_THREADS);
        const long long cap = (long long)c0->sm_count * 16;
        for (int pass = 0; pass < 4; ++pass) {
            {
                ProfScope ps(c0, 4);
                launch_diag_hist_b(pass, (unsigned)(pass < 2 ? blk : std::min(blk, cap)), st, dt, count, (int)blk);
            }
            k_diag_select_b<<<sel, MV_THREADS, 0, st>>>(dp, 1, pass, (int)sel);
            c0->launches += 1;
        }
        k_diag_median_final_b<<<(unsigned)((nm + 255) / 256), 256, 0, st>>>(dp, 1);
        c0->launches += 1;
    } else {
        {
            ProfScope ps(c0, 4);
            k_diag_mean_partial_b<<<(unsigned)blk, MV_THREADS, 0, st>>>(dt, count);
        }
        k_diag_mean_reduce_b<<<(unsigned)red, 64 * DIAG_RED_SL, 0, st>>>(dt, count);
        k_diag_pool_mean_final<<<(unsigned)((nm + 255) / 256), 256, 0, st>>>(dt, count, (int)nm, davg);
        c0->launches += 2;
    }
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(avg, davg, sizeof(double) * (size_t)nm, cudaMemcpyDeviceToHost, st));
    CK(cudaEventRecord(c0->ev_ready, st));
    for (int m = 1; m < count; ++m) CK(cudaStreamWaitEvent(ctxs[m]->stream, c0->ev_ready, 0));
    CK(cudaStreamSynchronize(st));
    return 0;
}

/* the per-diagonal statistic left resident by the last gcx_diag_stats / gcx_mcfs_batch (n doubles) */
extern "C" int gcx_diag_avg_read(gcx_ctx* c, double* avg) {
    CKR(need_map(c));
    if (!avg) return fail("gcx_diag_avg_read: null output");
    CK(cudaMemcpyAsync(avg, c->vavg, sizeof(double) * c->n, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    return 0;
}

// ---------------------------------------------------------------------------------------------------
// compaction (inner-seam remove_zeros)
// ---------------------------------------------------------------------------------------------------
extern "C" int gcx_compact_f64(gcx_ctx* c, const int32_t* kept, int64_t nk, double* host_out) {
    CKR(need_map(c));
    if (c->world != 1) return fail("gcx_compact_f64: single-rank maps only");
    if (nk < 1 || nk > c->n) return fail("gcx_compact_f64: bad kept count");
    const size_t b_idx = align256(sizeof(int) * (size_t)nk);
    const int64_t rows_per = std::max<int64_t>(1, std::min<int64_t>(nk, (int64_t)(256ll << 20) / (8 * nk)));   // <= 256 MB per slab
    CKR(ensure_scratch(c, b_idx + sizeof(double) * (size_t)rows_per * nk));
    int* didx = (int*)c->scratch;
    double* dout = (double*)((char*)c->scratch + b_idx);
    CK(cudaMemcpyAsync(didx, kept, sizeof(int) * (size_t)nk, cudaMemcpyHostToDevice, c->stream));
    for (int64_t r0 = 0; r0 < nk; r0 += rows_per) {
        const int rows = (int)std::min<int64_t>(rows_per, nk - r0);
        {
            ProfScope ps(c, 7);
            k_compact_f64<<<std::min(rows, c->sm_count * 8), 256, 0, c->stream>>>(c->A, c->ld, didx + r0, rows, didx, (int)nk, dout);
        }
        CK(cudaGetLastError());
        CK(cudaMemcpyAsync(host_out + r0 * nk, dout, sizeof(double) * (size_t)rows * nk, cudaMemcpyDeviceToHost, c->stream));
        CK(cudaStreamSynchronize(c->stream));
    }
    return 0;
}

// ---------------------------------------------------------------------------------------------------
// coarsening
// ---------------------------------------------------------------------------------------------------
extern "C" int gcx_downsample(gcx_ctx* c, int level, int method, float* host_out, int64_t out_n, float* minv, float* maxv) {
    CKR(need_map(c));
    if (c->world != 1) return fail("gcx_downsample: single-rank maps only");
    if (level < 1 || level > 128) return fail("gcx_downsample: level must be in [1, 128]");
    if (method < 0 || method > 2) return fail("gcx_downsample: method must be 0 (sum), 1 (mean) or 2 (max)");
    const int n = (int)c->n;
    const int64_t on = (n - 1 + level - 1) / level + 1;         // lib/ccmap.py:731-734
    if (out_n != on) return fail("gcx_downsample: output must be %lld x %lld", (long long)on, (long long)on);
    const bool padded = ((on - 1) * level - (n - 1)) != 0;      // :786-788: padding promotes the arithmetic to float64
    CKR(ensure_scratch(c, sizeof(float) * (size_t)on * on));
    float* dout = (float*)c->scratch;
    CKR(mm_begin(c));
    const int grid = (int)std::min<int64_t>((on * on + 255) / 256, (int64_t)c->sm_count * 16);
    {
        ProfScope ps(c, 7);
        if (padded) k_downsample<double><<<grid, 256, 0, c->stream>>>(c->A, c->ld, n, level, method, (int)on, dout, c->mm);
        else k_downsample<float><<<grid, 256, 0, c->stream>>>(c->A, c->ld, n, level, method, (int)on, dout, c->mm);
    }
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(host_out, dout, sizeof(float) * (size_t)on * on, cudaMemcpyDeviceToHost, c->stream));
    return mm_end(c, minv, maxv);
}

// ---------------------------------------------------------------------------------------------------
// transition probabilities + stationary distribution
// ---------------------------------------------------------------------------------------------------
#define EW_DISPATCH_B(KERNEL, FLAG, ...)                                                                            \
    switch (ew_variant()) {                                                                                        \
        case 1: KERNEL<8, 2, FLAG><<<c->sm_count * 2, MV_THREADS, 0, c->stream>>>(__VA_ARGS__); break;              \
        case 2: KERNEL<8, 3, FLAG><<<c->sm_count * 3, MV_THREADS, 0, c->stream>>>(__VA_ARGS__); break;              \
        case 3: KERNEL<4, 4, FLAG><<<c->sm_count * 4, MV_THREADS, 0, c->stream>>>(__VA_ARGS__); break;              \
        case 4: KERNEL<4, 6, FLAG><<<c->sm_count * 6, MV_THREADS, 0, c->stream>>>(__VA_ARGS__); break;              \
        case 5: KERNEL<2, 8, FLAG><<<c->sm_count * 8, MV_THREADS, 0, c->stream>>>(__VA_ARGS__); break;              \
        default: KERNEL<4, 3, FLAG><<<c->sm_count * 3, MV_THREADS, 0, c->stream>>>(__VA_ARGS__); break;             \
    }

extern "C" int gcx_transition_run(gcx_ctx* c, int in_place, float* minv, float* maxv) {
    CKR(need_map(c));
    CKR(fence_copy(c));
    if (!c->have_mask) return fail("gcx_transition_run: call gcx_set_mask first");
    CKR(run_row_stats(c));
    // sumr = A[m].sum() is float32 in the reference (lib/statDist.py:83)
    k_f64_to_f32<<<(int)((c->nv + 255) / 256), 256, 0, c->stream>>>((int)c->nv, c->vrowsum, c->csum32);
    CK(cudaGetLastError());
    c->launches += 1;
    float* out = c->A;
    if (!in_place) { CKR(ensure_out(c)); out = c->Out; }
    CKR(mm_begin(c));
    {
        ProfScope ps(c, 2);
        EW_DISPATCH_B(k_tp_pass, false, c->A, out, c->ld, (int)c->nloc, (int)(c->ld / 4), (int)c->n, (int)c->row0, c->csum32, c->keep, 0.f, c->mm)
    }
    CK(cudaGetLastError());
    float fill = 0.f, qmax = 0.f;
    CKR(mm_end(c, &fill, &qmax));
    if (fill != fill) return fail("gcx_transition_run: no non-zero entry in the kept rows");
    CKR(mm_begin(c));
    {
        ProfScope ps(c, 2);
        EW_DISPATCH_B(k_tp_pass, true, c->A, out, c->ld, (int)c->nloc, (int)(c->ld / 4), (int)c->n, (int)c->row0, c->csum32, c->keep, fill, c->mm)
    }
    CK(cudaGetLastError());
    if (in_place) { c->have_rowsum = false; c->sym_state = -1; }
    return mm_end(c, minv, maxv);
}

extern "C" int gcx_stationary_run(gcx_ctx* c, double tol, int max_iter, double* x, int* iterations, double* last_delta) {
    CKR(need_map(c));
    if (!c->have_mask) return fail("gcx_stationary_run: call gcx_set_mask first");
    if (max_iter < 1) return fail("gcx_stationary_run: max_iter must be positive");
    const int n = (int)c->n, n4 = (int)(c->ld / 4), nloc = (int)c->nloc;
    // m0: smallest non-zero entry of the kept block (:466-469)
    CKR(mm_begin(c));
    {
        ProfScope ps(c, 3);
        EW_DISPATCH(k_block_min, c->A, c->ld, nloc, n4, n, (int)c->row0, c->keep, c->mm)
    }
    CK(cudaGetLastError());
    float m0 = 0.f, mx = 0.f;
    CKR(mm_end(c, &m0, &mx));
    if (m0 != m0) return fail("gcx_stationary_run: the kept block has no non-zero entry");
    const int nch = (n4 + MV_THREADS - 1) / MV_THREADS;
    double* xv = c->kr.x;
    double* yv = c->vq;
    constexpr int TKG = 4, TSTAGES = 6, TOCC = 2;                 // 16 KB per stage, 96 KB per CTA, 2 CTAs per SM (like the dense TMA product)
    const size_t smem = (size_t)TSTAGES * TKG * MV_THREADS * sizeof(float4);
    int& mvt_occ = c->mvt_occ;                 // per context: function attributes and occupancy belong to the device
    if (mvt_occ < 0) {
        CK(cudaFuncSetAttribute(k_matvec_t_tma<TKG, TSTAGES, TOCC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        int occ = 0;
        CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_matvec_t_tma<TKG, TSTAGES, TOCC>, TMA_THREADS, smem));
        mvt_occ = std::min(occ, TOCC);
    }
    const bool use_tma = c->sd_variant == 1 && mvt_occ >= 1;
    // LDG variant: (column chunks x row chunks) grid, ~6 CTAs per SM, chunk length a multiple of the unroll depth
    int nrc = std::max(1, (c->sm_count * 6 + nch - 1) / nch);
    int rpc = (nloc + nrc - 1) / nrc;
    rpc = std::max(MVT_UNR, (rpc + MVT_UNR - 1) / MVT_UNR * MVT_UNR);
    nrc = (nloc + rpc - 1) / rpc;
    // TMA variant: persistent grid over the strip-major unit list; tabulate every CTA's first strip and every strip's CTA range
    const int G = c->sm_count * std::max(1, mvt_occ);
    MvtTab tab{};
    size_t part_bytes = sizeof(double) * (size_t)nrc * c->ld, tab_bytes = 0;
    std::vector<int> ht;
    const int nst = (nch + TKG - 1) / TKG;
    if (use_tma) {
        const long long U = (long long)nst * nloc;
        std::vector<int> gf((size_t)G, 0), lo((size_t)nst, G), hi((size_t)nst, -1);
        int slots = 1;
        for (int g = 0; g < G; ++g) {
            const long long u0 = (long long)g * U / G, u1 = (long long)(g + 1) * U / G;
            if (u1 <= u0) continue;
            const int s0 = (int)(u0 / nloc), s1 = (int)((u1 - 1) / nloc);
            gf[g] = s0;
            slots = std::max(slots, s1 - s0 + 1);
            for (int k = s0; k <= s1; ++k) { lo[k] = std::min(lo[k], g); hi[k] = std::max(hi[k], g); }
        }
        ht = gf;
        ht.insert(ht.end(), lo.begin(), lo.end());
        ht.insert(ht.end(), hi.begin(), hi.end());
        tab.slots = slots;
        part_bytes = sizeof(double) * (size_t)G * slots * TKG * 1024;
        tab_bytes = align256(sizeof(int) * ht.size());
    }
    CKR(ensure_scratch(c, align256(part_bytes) + tab_bytes));
    double* part = (double*)c->scratch;
    if (use_tma) {
        int* dt = (int*)((char*)c->scratch + align256(part_bytes));
        CK(cudaMemcpyAsync(dt, ht.data(), sizeof(int) * ht.size(), cudaMemcpyHostToDevice, c->stream));
        CK(cudaStreamSynchronize(c->stream));          // ht is a local
        tab.cta_gf = dt;
        tab.strip_lo = dt + G;
        tab.strip_hi = dt + G + nst;
    }
    k_sd_init<<<EP_CL, EP_THREADS, 0, c->stream>>>(c->sd_state, (int)c->nv, n, c->keep, xv, tol, max_iter);
    CK(cudaGetLastError());
    c->launches += 1;
    c->last_sym = false;
    c->have_scale = false;           // the KR vector slab is reused
    auto step = [&]() -> int {
        {
            ProfScope ps(c, 0);
            if (use_tma)
                k_matvec_t_tma<TKG, TSTAGES, TOCC><<<G, TMA_THREADS, smem, c->stream>>>(c->A, c->ld, nloc, n4, (int)c->row0, xv, m0, part, tab,
                                                                                      &c->sd_state->done);
            else
                k_matvec_t<<<dim3(nch, nrc), MV_THREADS, 0, c->stream>>>(c->A, c->ld, nloc, n4, (int)c->row0, rpc, xv, m0, part, &c->sd_state->done);
        }
        CK(cudaGetLastError());
        if (use_tma)
            k_sd_fold_t<TKG><<<(int)((c->nv + 255) / 256), 256, 0, c->stream>>>(c->sd_state, (int)c->nv, n, c->keep, part, tab, yv);
        else
            k_sd_fold<<<(int)((c->nv + 255) / 256), 256, 0, c->stream>>>(c->sd_state, (int)c->nv, n, nrc, c->ld, c->keep, part, yv);
        CK(cudaGetLastError());
        c->launches += 1;
        CKR(allreduce_f64(c, yv, (size_t)c->n));
        {
            ProfScope ps(c, 6);
            k_sd_step<<<EP_CL, EP_THREADS, 0, c->stream>>>(c->sd_state, n, c->keep, yv, xv);
        }
        CK(cudaGetLastError());
        return 0;
    };
    SdState fin;
    CKR(iterate<SdState>(c, c->sd_state, max_iter + 4, step, &fin));
    if (x) CK(cudaMemcpyAsync(x, xv, sizeof(double) * n, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    if (iterations) *iterations = fin.iters;
    if (last_delta) *last_delta = fin.delta;
    return 0;
}

// ---------------------------------------------------------------------------------------------------
// correlation matrix
// ---------------------------------------------------------------------------------------------------
static int blas_ready(gcx_ctx* c) {
    CKR(cublas_load());
    if (!c->blas) {
        BK(g_cublas.Create(&c->blas));
        BK(g_cublas.SetStream(c->blas, c->stream));
    }
    return 0;
}

// row-major P = L R^T for row-major m x m factors (column-major view: P^T = R_cm^T L_cm)
static int gemm_nt(gcx_ctx* c, int m, const double* L, const double* R, double* P) {
    const double one = 1.0, zero = 0.0;
    ProfScope ps(c, 7);
    BK(g_cublas.Dgemm(c->blas, CUBLAS_OP_T_, CUBLAS_OP_N_, m, m, m, &one, R, m, L, m, &zero, P, m));
    return 0;
}

struct CorrBufs {
    double *X0, *I, *N, *Sx, *Sxy, *var;
    float* Nf;               // counts from the bf16 tensor-core product (corr_variant 1)
    unsigned short* Ib;      // bf16 copy of the indicator as an mb x mb matrix, mb = m rounded up to 64 and the pad zero, so that every
    int mb;                  // dimension and pitch of the counts product is aligned for cuBLAS' tensor-core kernels; Nf has pitch mb too
    signed char* S8;         // up to 9 seven-bit slices of X0 as int8 (mb x mb each, pad zero)
    int* rhi;                // per-row exponents of the sliced float32 path
    signed char* I8;         // int8 indicator (mb x mb)
    int* T32;                // one slice product (int32, mb x mb)
    int* kept;
};

static int corr_bufs(gcx_ctx* c, int64_t m, CorrBufs* B) {
    const size_t mm = align256(sizeof(double) * (size_t)m * m), bv = align256(sizeof(double) * (size_t)m), bi = align256(sizeof(int) * (size_t)m);
    const int64_t mb = round_up(m, 256);       // every tile of k_i8gemm_nt (128 x 256 x 128) is full; also satisfies cuBLAS' 64-element alignment
    const size_t bb = align256(sizeof(unsigned short) * (size_t)mb * mb);
    const size_t nn = std::max(mm, align256(sizeof(float) * (size_t)mb * mb));      // N (float64 m x m) doubles as Nf (float32 mb x mb)
    const size_t tt = align256(sizeof(int) * (size_t)mb * mb), b8 = align256((size_t)mb * mb);
    const size_t need = 4 * mm + nn + bb + 10 * b8 + tt + bv + 2 * bi;
    size_t free_b = 0, total_b = 0;
    CK(cudaMemGetInfo(&free_b, &total_b));
    if (need > free_b + c->scratch_bytes)
        return fail("correlation of %lld bins needs %.1f GB of workspace, %.1f GB are free", (long long)m, need / 1e9, (free_b + c->scratch_bytes) / 1e9);
    CKR(ensure_scratch(c, need));
    char* p = (char*)c->scratch;
    B->X0 = (double*)p; p += mm;
    B->I = (double*)p; p += mm;
    B->N = (double*)p; B->Nf = (float*)p; p += nn;
    B->Sx = (double*)p; p += mm;
    B->Sxy = (double*)p; p += mm;
    B->Ib = (unsigned short*)p; p += bb;
    B->mb = (int)mb;
    B->S8 = (signed char*)p; p += 9 * b8;
    B->I8 = (signed char*)p; p += b8;
    B->T32 = (int*)p; p += tt;
    CK(cudaMemsetAsync(B->Ib, 0, bb + 10 * b8, c->stream));   // indicator and slice pads
    B->var = (double*)p; p += bv;
    B->kept = (int*)p; p += bi;
    B->rhi = (int*)p;
    return 0;
}

// ---- this library's int8 tensor-core product: row-major T32 = L R^T on k_i8gemm_nt (tcgen05 + TMEM + TMA) ----
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                    const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static PFN_encodeTiled g_encode_tiled = nullptr;
static int i8_tensor_map(CUtensorMap* tm, const signed char* base, int mb, int box_rows) {
    if (!g_encode_tiled) {
        std::lock_guard<std::mutex> lock(g_loader_mutex);
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
        if (!fn || qres != cudaDriverEntryPointSuccess) return fail("cuTensorMapEncodeTiled is not available in this driver");
        g_encode_tiled = (PFN_encodeTiled)fn;
    }
    const cuuint64_t dims[2] = {(cuuint64_t)mb, (cuuint64_t)mb};          // innermost first: K (bytes), rows
    const cuuint64_t strides[1] = {(cuuint64_t)mb};                        // bytes between rows
    const cuuint32_t box[2] = {(cuuint32_t)I8_BK, (cuuint32_t)box_rows};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = g_encode_tiled(tm, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, (void*)base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                      CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail("cuTensorMapEncodeTiled failed with status %d", (int)r);
    return 0;
}

template <int BN, int STAGES>
static int i8gemm_launch(gcx_ctx* c, const signed char* L, const signed char* R, int* T32, int mb, int symmetric, unsigned int bit) {
    constexpr int smem = i8_smem_bytes(BN, STAGES);
    if (!(c->i8_cfg_mask & bit)) {
        CK(cudaFuncSetAttribute(k_i8gemm_nt<BN, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        c->i8_cfg_mask |= bit;
    }
    CUtensorMap tmL, tmR;
    CKR(i8_tensor_map(&tmL, L, mb, I8_BM));
    CKR(i8_tensor_map(&tmR, R, mb, BN));
    {
        ProfScope ps(c, 7);
        k_i8gemm_nt<BN, STAGES><<<c->sm_count, I8_THREADS, smem, c->stream>>>(tmL, tmR, T32, mb, symmetric);
    }
    CK(cudaGetLastError());
    return 0;
}

static int i8gemm_launch_pair(gcx_ctx* c, const signed char* L, const signed char* R, int* T32, int mb, int symmetric, const I8Epilogue& E) {
    if (!(c->i8_cfg_mask & 16u)) {
        CK(cudaFuncSetAttribute(k_i8gemm_nt_pair, cudaFuncAttributeMaxDynamicSharedMemorySize, I8P_SMEM));
        c->i8_cfg_mask |= 16u;
    }
    CUtensorMap tmL, tmR;
    CKR(i8_tensor_map(&tmL, L, mb, 128));
    CKR(i8_tensor_map(&tmR, R, mb, 128));
    {
        ProfScope ps(c, 7);
        k_i8gemm_nt_pair<<<c->sm_count & ~1, I8_THREADS, I8P_SMEM, c->stream>>>(tmL, tmR, T32, mb, symmetric, E);
    }
    CK(cudaGetLastError());
    return 0;
}

static int i8_variant() {
    // default: the CTA-pair kernel (cta_group::2, 256 x 256 tiles; 3.0-3.5 POP/s); 0: one CTA per 128 x 256 tile (2.7 POP/s); 1-3: sweeps
    static const int variant = [] { const char* e = getenv("GCX_I8_VARIANT"); return e ? atoi(e) : 4; }();
    return variant;
}

// product with the accumulation fused into the epilogue (pair kernel only; the caller checks i8_variant() == 4)
// The fused accumulation is correct (tests) but SLOWER today (72.8 ms vs 63.3 ms per correlation at n = 24,926): a thread owns one row of
// the tile, so its float64 read-modify-writes touch 32 different rows per instruction (8 useful bytes per 32-byte sector) and the
// epilogue stalls the MMA pipeline.  Off by default (GCX_I8_FUSED=1 enables it) until the epilogue stages the tile through shared
// memory for coalesced row segments.
static bool i8_fused_epilogue() {
    static const bool on = [] { const char* e = getenv("GCX_I8_FUSED"); return e && atoi(e) != 0; }();
    return on && i8_variant() == 4;
}

static int i8gemm_nt_fused(gcx_ctx* c, const signed char* L, const signed char* R, int mb, int symmetric, const I8Epilogue& E) {
    if (mb % 256 != 0) return fail("i8gemm_nt: the padded size %d is not a multiple of 256", mb);
    return i8gemm_launch_pair(c, L, R, nullptr, mb, symmetric, E);
}

static int i8gemm_nt(gcx_ctx* c, const signed char* L, const signed char* R, int* T32, int mb, int symmetric) {
    if (mb % 256 != 0) return fail("i8gemm_nt: the padded size %d is not a multiple of 256", mb);
    const I8Epilogue store{0, 0, nullptr, nullptr, 0, 0, 0};
    switch (i8_variant()) {
        case 4: return i8gemm_launch_pair(c, L, R, T32, mb, symmetric, store);
        case 1: return i8gemm_launch<256, 3>(c, L, R, T32, mb, symmetric, 2u);
        case 2: return i8gemm_launch<128, 6>(c, L, R, T32, mb, symmetric, 4u);
        case 3: return i8gemm_launch<128, 4>(c, L, R, T32, mb, symmetric, 8u);
        default: return i8gemm_launch<256, 4>(c, L, R, T32, mb, symmetric, 1u);
    }
}

// the three products and the row variances, from X0 / I.  corr_variant 1 (default): counts on the bf16 tensor cores + SYRK for
// the symmetric sum; 0: three float64 GEMMs.  Returns which count buffer the epilogue must read.
static int corr_core(gcx_ctx* c, int m, const CorrBufs& B, int* counts_mode) {
    CKR(blas_ready(c));
    *counts_mode = c->corr_variant == 1 ? 1 : 0;          // 0: float64 N, 1: float32 Nf (bf16 product), 2: int32 in Nf (int8 product)
    if (c->corr_variant == 1) {
        const float one = 1.f, zero = 0.f;
        const double done = 1.0, dzero = 0.0;
        auto counts_bf16 = [&]() -> int {
            ProfScope ps(c, 7);      // row-major N = I I^T: column-major (I_cm)^T I_cm, bf16 x bf16 -> float32
            BK(g_cublas.GemmEx(c->blas, CUBLAS_OP_T_, CUBLAS_OP_N_, B.mb, B.mb, B.mb, &one, B.Ib, 14 /* CUDA_R_16BF */, B.mb, B.Ib, 14, B.mb, &zero,
                               B.Nf, 0 /* CUDA_R_32F */, B.mb, 68 /* CUBLAS_COMPUTE_32F */, -1 /* CUBLAS_GEMM_DEFAULT */));
            return 0;
        };
        auto counts_int8 = [&]() -> int {   // the sliced routes have the int8 indicator at hand: same product at twice the rate, int32 counts
            const int i1 = 1, i0 = 0;
            *counts_mode = 2;
            if (c->corr_own_gemm) return i8gemm_nt(c, B.I8, B.I8, (int*)B.Nf, B.mb, 1);      // symmetric: the consumers read j <= i only
            ProfScope ps(c, 7);
            BK(g_cublas.GemmEx(c->blas, CUBLAS_OP_T_, CUBLAS_OP_N_, B.mb, B.mb, B.mb, &i1, B.I8, 3 /* CUDA_R_8I */, B.mb, B.I8, 3, B.mb, &i0, B.Nf,
                               10 /* CUDA_R_32I */, B.mb, 72 /* CUBLAS_COMPUTE_32I */, -1));
            return 0;
        };
        // integer-valued map: Sx and Sxy from exact int8 slice products (int32 accumulation) instead of the float64 GEMM / SYRK
        int ns = 0;
        if (c->corr_int_path && 16129LL * m < (1LL << 31)) {
            CK(cudaMemsetAsync(c->csum_dev, 0, sizeof(unsigned long long), c->stream));
            k_corr_intcheck<<<c->sm_count * 8, 256, 0, c->stream>>>(B.X0, (long long)m * m, (unsigned int*)c->csum_dev);
            CK(cudaGetLastError());
            c->launches += 1;
            unsigned int f[2] = {1u, 0u};
            CKR(read_small(c, c->csum_dev, sizeof f, f));
            if (!(f[0] & 1u)) {
                int bits = 0;
                while (bits < 32 && (f[1] >> bits)) ++bits;
                ns = std::max(1, (bits + 6) / 7);
            }
        }
        if (ns > 0) {
            const int ione = 1, izero = 0;
            const size_t plane = (size_t)B.mb * B.mb;
            k_corr_slices8<<<c->sm_count * 8, 256, 0, c->stream>>>(B.X0, B.I, m, B.mb, ns, B.S8, B.I8);
            CK(cudaGetLastError());
            c->launches += 1;
            auto imma = [&](const signed char* L, const signed char* R) -> int {        // row-major T32 = L R^T
                if (c->corr_own_gemm) return i8gemm_nt(c, L, R, B.T32, B.mb, L == R ? 1 : 0);     // S_a S_a^T: lower triangle only
                ProfScope ps(c, 7);
                return g_cublas.GemmEx(c->blas, CUBLAS_OP_T_, CUBLAS_OP_N_, B.mb, B.mb, B.mb, &ione, R, 3 /* CUDA_R_8I */, B.mb, L, 3, B.mb, &izero,
                                       B.T32, 10 /* CUDA_R_32I */, B.mb, 72 /* CUBLAS_COMPUTE_32I */, -1);
            };
            auto accum = [&](double scale, int first, double* dst) -> int {
                k_corr_accum32<<<c->sm_count * 8, 256, 0, c->stream>>>(B.T32, m, B.mb, scale, first, dst);
                CK(cudaGetLastError());
                c->launches += 1;
                return 0;
            };
            const bool fused = c->corr_own_gemm && i8_fused_epilogue();      // accumulate in the product's epilogue: no int32 round trip
            const int st = fused ? i8gemm_nt_fused(c, B.S8, B.I8, B.mb, 0, I8Epilogue{1, m, B.Sx, nullptr, 0, 1, 0}) : imma(B.S8, B.I8);
            if (st != 0 && fused) return st;
            if (st != 0) {
                ns = 0;                       // this cuBLAS build has no int8 path for the shape: float64 products from now on
                c->corr_int_path = 0;
            } else {
                if (!fused) CKR(accum(1.0, 1, B.Sx));
                CKR(counts_int8());
                for (int t = 1; t < ns; ++t) {
                    if (fused) {
                        CKR(i8gemm_nt_fused(c, B.S8 + t * plane, B.I8, B.mb, 0, I8Epilogue{1, m, B.Sx, nullptr, -7 * t, 0, 0}));
                    } else {
                        BK(imma(B.S8 + t * plane, B.I8));
                        CKR(accum((double)(1ull << (7 * t)), 0, B.Sx));
                    }
                }
                // S_b S_a^T is the transpose of S_a S_b^T: multiply a <= b only, add T + T^T on the lower triangle (exponents all 0)
                CK(cudaMemsetAsync(B.rhi, 0, sizeof(int) * (size_t)m, c->stream));
                const dim3 tg((m + 31) / 32, (m + 31) / 32);
                bool first = true;
                for (int a = 0; a < ns; ++a)
                    for (int b = a; b < ns; ++b) {
                        if (fused && a == b) {
                            // S_a S_a^T: lower triangle, accumulated into Sxy by the epilogue (k_corr_faccum_sxy with sym = 0)
                            CKR(i8gemm_nt_fused(c, B.S8 + a * plane, B.S8 + a * plane, B.mb, 1, I8Epilogue{1, m, B.Sxy, B.rhi, -7 * (a + b), first ? 1 : 0, 1}));
                        } else {
                            BK(imma(B.S8 + a * plane, B.S8 + b * plane));
                            k_corr_faccum_sxy<<<tg, 256, 0, c->stream>>>(B.T32, m, B.mb, B.rhi, -7 * (a + b), a != b, first ? 1 : 0, B.Sxy);
                            CK(cudaGetLastError());
                            c->launches += 1;
                        }
                        first = false;
                    }
            }
        }
        c->corr_last_int = ns;
        if (ns == 0) {
            // non-integer map: Sx still exact on the int8 units when every entry is a float32 number whose row fits 63 bits
            int K = 0;
            if (c->corr_int_path && 127LL * m < (1LL << 31)) {
                CK(cudaMemsetAsync(c->csum_dev, 0, sizeof(unsigned long long), c->stream));
                k_corr_f32check<<<c->sm_count * 8, 256, 0, c->stream>>>(B.X0, m, B.rhi, (unsigned int*)c->csum_dev);
                CK(cudaGetLastError());
                c->launches += 1;
                unsigned int f[2] = {1u, 0u};
                CKR(read_small(c, c->csum_dev, sizeof f, f));
                if (!(f[0] & 1u) && f[1] <= 63u) K = std::max(1, ((int)f[1] + 6) / 7);
            }
            if (K > 0) {
                const int ione = 1, izero = 0;
                const size_t plane = (size_t)B.mb * B.mb;
                k_corr_fslices8<<<c->sm_count * 8, 256, 0, c->stream>>>(B.X0, B.I, m, B.mb, K, B.rhi, B.S8, B.I8);
                CK(cudaGetLastError());
                c->launches += 1;
                CKR(counts_int8());
                const bool fusedf = c->corr_own_gemm && i8_fused_epilogue();
                for (int t = 0; t < K; ++t) {
                    if (fusedf) {          // Sx += 2^(rhi_i - 7 (t + 1)) S_t I^T in the product's epilogue
                        CKR(i8gemm_nt_fused(c, B.S8 + t * plane, B.I8, B.mb, 0, I8Epilogue{1, m, B.Sx, B.rhi, 7 * (t + 1), t == 0 ? 1 : 0, 0}));
                        continue;
                    }
                    if (c->corr_own_gemm) {
                        CKR(i8gemm_nt(c, B.S8 + t * plane, B.I8, B.T32, B.mb, 0));
                    } else {
                        ProfScope ps(c, 7);      // row-major T32 = S_t I^T
                        BK(g_cublas.GemmEx(c->blas, CUBLAS_OP_T_, CUBLAS_OP_N_, B.mb, B.mb, B.mb, &ione, B.I8, 3 /* CUDA_R_8I */, B.mb,
                                           B.S8 + t * plane, 3, B.mb, &izero, B.T32, 10 /* CUDA_R_32I */, B.mb, 72 /* CUBLAS_COMPUTE_32I */, -1));
                    }
                    k_corr_faccum<<<c->sm_count * 8, 256, 0, c->stream>>>(B.T32, m, B.mb, B.rhi, 7 * (t + 1), t == 0, B.Sx);
                    CK(cudaGetLastError());
                    c->launches += 1;
                }
                c->corr_last_int = -K;
            } else {
                CKR(counts_bf16());
                CKR(gemm_nt(c, m, B.X0, B.I, B.Sx));
            }
            if (K > 0 && K <= 8 && 16129LL * m < (1LL << 31)) {
                // Sxy exactly from K (K + 1) / 2 int8 products (<= 36, ~10 ms each with the accumulation at n = 24,926; the SYRK takes 434 ms)
                const int ione = 1, izero = 0;
                const size_t plane = (size_t)B.mb * B.mb;
                const dim3 tg((m + 31) / 32, (m + 31) / 32);
                bool first = true;
                for (int a = 0; a < K; ++a)
                    for (int b = a; b < K; ++b) {
                        if (c->corr_own_gemm && i8_fused_epilogue() && a == b) {
                            CKR(i8gemm_nt_fused(c, B.S8 + a * plane, B.S8 + a * plane, B.mb, 1, I8Epilogue{1, m, B.Sxy, B.rhi, 7 * (a + b + 2), first ? 1 : 0, 1}));
                            first = false;
                            continue;
                        }
                        if (c->corr_own_gemm) {
                            CKR(i8gemm_nt(c, B.S8 + a * plane, B.S8 + b * plane, B.T32, B.mb, a == b ? 1 : 0));
                        } else {
                            ProfScope ps(c, 7);      // row-major T32 = S_a S_b^T
                            BK(g_cublas.GemmEx(c->blas, CUBLAS_OP_T_, CUBLAS_OP_N_, B.mb, B.mb, B.mb, &ione, B.S8 + b * plane, 3, B.mb,
                                               B.S8 + a * plane, 3, B.mb, &izero, B.T32, 10, B.mb, 72, -1));
                        }
                        k_corr_faccum_sxy<<<tg, 256, 0, c->stream>>>(B.T32, m, B.mb, B.rhi, 7 * (a + b + 2), a != b, first ? 1 : 0, B.Sxy);
                        CK(cudaGetLastError());
                        c->launches += 1;
                        first = false;
                    }
            } else {
                ProfScope ps(c, 7);      // row-major lower triangle of X0 X0^T = column-major upper triangle of (X_cm)^T X_cm
                BK(g_cublas.Dsyrk(c->blas, 1 /* CUBLAS_FILL_MODE_UPPER */, CUBLAS_OP_T_, m, m, &done, B.X0, m, &dzero, B.Sxy, m));
            }
        }
    } else {
        c->corr_last_int = 0;
        CKR(gemm_nt(c, m, B.I, B.I, B.N));
        CKR(gemm_nt(c, m, B.X0, B.I, B.Sx));
        CKR(gemm_nt(c, m, B.X0, B.X0, B.Sxy));
    }
    {
        static const int sv = [] { const char* e = getenv("GCX_CORR_SELFVAR"); return e ? atoi(e) : 1; }();
        if (sv == 1) {
            const int smem = SV_WARPS * SV_STAGES * 2 * 32 * SV_PITCH * (int)sizeof(double);
            if (!c->selfvar_ready) {
                CK(cudaFuncSetAttribute(k_corr_selfvar_tiled, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
                c->selfvar_ready = true;
            }
            k_corr_selfvar_tiled<<<(m + SV_WARPS * 32 - 1) / (SV_WARPS * 32), SV_WARPS * 32, smem, c->stream>>>(B.X0, B.I, m, B.var);
        } else {
            k_corr_selfvar<<<(m + 127) / 128, 128, 0, c->stream>>>(B.X0, B.I, m, B.var);
        }
    }
    CK(cudaGetLastError());
    c->launches += 1;
    return 0;
}

extern "C" int gcx_corr_run(gcx_ctx* c, int has_mask, double maskvalue, int logspace, int has_vmin, float vmin, int has_vmax, float vmax) {
    CKR(need_map(c));
    CKR(fence_copy(c));
    if (c->world != 1) return fail("gcx_corr_run: single-rank maps only");
    if (!c->have_mask) return fail("gcx_corr_run: call gcx_set_mask first");
    const int n = (int)c->n;
    std::vector<uint8_t> keep((size_t)n);
    CK(cudaMemcpyAsync(keep.data(), c->keep, (size_t)n, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    std::vector<int> kept;
    for (int i = 0; i < n; ++i)
        if (keep[i]) kept.push_back(i);
    const int m = (int)kept.size();
    if (m < 1) return fail("gcx_corr_run: no bin with data");
    CorrBufs B{};
    CKR(corr_bufs(c, m, &B));
    CK(cudaMemcpyAsync(B.kept, kept.data(), sizeof(int) * (size_t)m, cudaMemcpyHostToDevice, c->stream));
    CK(cudaStreamSynchronize(c->stream));          // `kept` is a local
    CorrPrep P{};
    P.has_mask = has_mask; P.logspace = logspace; P.has_vmin = has_vmin; P.has_vmax = has_vmax;
    P.maskf = (float)maskvalue;                      // the wrapper rounds the mask value to float32 (lib/_corrMatrixCore.pyx:88)
    P.mask = (double)P.maskf;
    P.vmin = vmin; P.vmax = vmax;
    {
        ProfScope ps(c, 7);
        k_corr_prep<<<dim3((m + 31) / 32, (m + 31) / 32), 256, 0, c->stream>>>(c->A, c->ld, B.kept, m, P, B.X0, B.I, B.Ib, B.mb);
    }
    CK(cudaGetLastError());
    int cm = 0;
    CKR(corr_core(c, m, B, &cm));
    const bool nf = cm != 0;
    CKR(ensure_out(c));
    CK(cudaMemsetAsync(c->Out, 0, sizeof(float) * (size_t)c->nloc * c->ld, c->stream));
    {
        ProfScope ps(c, 7);
        k_corr_finish<true><<<dim3((m + 31) / 32, (m + 31) / 32), 256, 0, c->stream>>>(nf ? nullptr : B.N, nf ? B.Nf : nullptr, B.Sx, B.Sxy, B.mb, cm == 2, B.var, m, B.kept, c->Out, c->ld, nullptr, 0);
    }
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(c->stream));
    return 0;
}

/* measurement / test hook for the library's own int8 tensor-core product (k_i8gemm_nt): T = L R^T for host int8 matrices (mb x mb,
 * row-major, mb a multiple of 256); T_host may be NULL (timing only); reps >= 1 launches, ms = mean device time of one */
extern "C" int gcx_debug_i8gemm(gcx_ctx* c, const signed char* L, const signed char* R, int mb, int32_t* T_host, int symmetric, int reps, double* ms) {
    if (!c) return fail("null context");
    CKR(set_dev(c));
    if (mb < I8_BN || mb % I8_BN) return fail("gcx_debug_i8gemm: mb must be a positive multiple of %d", I8_BN);
    const size_t plane = (size_t)mb * mb;
    CKR(ensure_scratch(c, align256(plane) * 2 + sizeof(int) * plane));
    signed char* dL = (signed char*)c->scratch;
    signed char* dR = dL + align256(plane);
    int* dT = (int*)(dR + align256(plane));
    if (L) CK(cudaMemcpyAsync(dL, L, plane, cudaMemcpyHostToDevice, c->stream));
    else CK(cudaMemsetAsync(dL, 1, plane, c->stream));
    if (R) CK(cudaMemcpyAsync(dR, R, plane, cudaMemcpyHostToDevice, c->stream));
    else CK(cudaMemsetAsync(dR, 1, plane, c->stream));
    CK(cudaMemsetAsync(dT, 0xff, sizeof(int) * plane, c->stream));
    CKR(i8gemm_nt(c, dL, dR, dT, mb, symmetric));          // warm
    CK(cudaEventRecord(c->run_a, c->stream));
    for (int i = 0; i < std::max(1, reps); ++i) CKR(i8gemm_nt(c, dL, dR, dT, mb, symmetric));
    CK(cudaEventRecord(c->run_b, c->stream));
    if (T_host) CK(cudaMemcpyAsync(T_host, dT, sizeof(int) * plane, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    float t = 0;
    CK(cudaEventElapsedTime(&t, c->run_a, c->run_b));
    if (ms) *ms = t / std::max(1, reps);
    return 0;
}

extern "C" int gcx_corr_last_route(gcx_ctx* c, int* slices, int* sxy_sliced) {
    if (!c || !slices) return fail("gcx_corr_last_route: null argument");
    *slices = c->corr_last_int;
    if (sxy_sliced) *sxy_sliced = c->corr_last_int > 0 || (c->corr_last_int < 0 && -c->corr_last_int <= 8);
    return 0;
}

extern "C" int gcx_corr_matrix_f64(gcx_ctx* c, const double* in, int64_t m64, int has_mask, double maskvalue, int covariance, double* out) {
    if (!c) return fail("gcx_corr_matrix_f64: null context");
    CKR(set_dev(c));
    if (!in || !out || m64 < 1 || m64 > 0x7fffffff) return fail("gcx_corr_matrix_f64: bad arguments");
    const int m = (int)m64;
    CorrBufs B{};
    CKR(corr_bufs(c, m, &B));
    const size_t bytes = sizeof(double) * (size_t)m * m;
    CK(cudaMemcpyAsync(B.X0, in, bytes, cudaMemcpyHostToDevice, c->stream));
    const float maskf = (float)maskvalue;            // lib/_corrMatrixCore.pyx:88
    k_corr_prep_f64<<<c->sm_count * 8, 256, 0, c->stream>>>(B.X0, B.I, B.Ib, m, B.mb, has_mask, (double)maskf);
    CK(cudaGetLastError());
    c->launches += 1;
    int cm = 0;
    CKR(corr_core(c, m, B, &cm));
    const bool nf = cm != 0;
    // the result overwrites the indicator buffer (no longer needed once the products are done)
    CK(cudaMemsetAsync(B.I, 0, bytes, c->stream));
    k_corr_finish<false><<<dim3((m + 31) / 32, (m + 31) / 32), 256, 0, c->stream>>>(nf ? nullptr : B.N, nf ? B.Nf : nullptr, B.Sx, B.Sxy, B.mb, cm == 2, B.var, m, nullptr, nullptr, 0, B.I, covariance);
    CK(cudaGetLastError());
    c->launches += 1;
    CK(cudaMemcpyAsync(out, B.I, bytes, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    return 0;
}

// ---------------------------------------------------------------------------------------------------
// measurement
// ---------------------------------------------------------------------------------------------------
extern "C" int gcx_profile(gcx_ctx* c, int enable) {
    CKR(set_dev(c));
    CKR(prof_fold(c));
    c->profile = enable != 0;
    return 0;
}

extern "C" int gcx_profile_read(gcx_ctx* c, int reset, int64_t launches[GCX_NCLASS], double ms[GCX_NCLASS]) {
    CKR(set_dev(c));
    CKR(prof_fold(c));
    for (int i = 0; i < GCX_NCLASS; ++i) {
        launches[i] = c->prof_launches[i];
        ms[i] = c->prof_ms[i];
        if (reset) { c->prof_launches[i] = 0; c->prof_ms[i] = 0; }
    }
    return 0;
}

extern "C" int gcx_timer_start(gcx_ctx* c) {
    CKR(set_dev(c));
    CK(cudaStreamSynchronize(c->stream));
    CK(cudaEventRecord(c->tm_a, c->stream));
    return 0;
}

extern "C" int gcx_timer_stop(gcx_ctx* c, double* ms) {
    CKR(set_dev(c));
    CK(cudaEventRecord(c->tm_b, c->stream));
    CK(cudaEventSynchronize(c->tm_b));
    float f = 0;
    CK(cudaEventElapsedTime(&f, c->tm_a, c->tm_b));
    *ms = f;
    return 0;
}

extern "C" int gcx_map_invalidate(gcx_ctx* c) {
    c->have_rowsum = false;
    return 0;
}

/* diagnostics: [G][16] int64 per CTA of launch GCX_SYM_DEBUG_LAUNCH (default 20) of the deferred-scalar IC pass -- globaltimer ns at
 * 0 entry, 1 first tiles issued, 2 vector phase done, 3 past the grid barrier, 4 product starts, 5 static share done, 6 all done;
 * 7 = dynamic chunks taken */
extern "C" int gcx_debug_sym_stamps(gcx_ctx* c, int64_t* stamps, int capacity, int* count) {
    CKR(need_map(c));
    if (!c->sym_stamps) return fail("phase stamps are off (set GCX_SYM_DEBUG=1 before the first symmetric pass)");
    const int G = c->sym_grid;
    if (capacity < 16 * G) return fail("capacity < 16 x grid (%d)", G);
    CK(cudaStreamSynchronize(c->stream));
    CK(cudaMemcpy(stamps, c->sym_stamps, sizeof(long long) * 16 * (size_t)G, cudaMemcpyDeviceToHost));
    *count = G;
    return 0;
}

extern "C" int gcx_debug_sym_cycles(gcx_ctx* c, int64_t* cycles, int capacity, int* count) {
    CKR(need_map(c));
    if (!c->sym_dbg) return fail("per-CTA cycle recording is off (set GCX_SYM_DEBUG=1 before the first symmetric pass)");
    const int G = c->sym_grid;
    if (capacity < G) return fail("capacity < grid (%d)", G);
    CK(cudaStreamSynchronize(c->stream));
    CK(cudaMemcpy(cycles, c->sym_dbg, sizeof(long long) * (size_t)G, cudaMemcpyDeviceToHost));
    *count = G;
    return 0;
}

extern "C" int gcx_pass_bytes(gcx_ctx* c, int64_t* bytes, int* symmetric) {
    CKR(need_map(c));
    if (symmetric) *symmetric = c->last_sym ? 1 : 0;
    *bytes = c->last_sym ? sym_bytes_per_pass(c) : (int64_t)4 * c->nloc * c->n;
    return 0;
}

extern "C" int gcx_last_run_products(gcx_ctx* c, int* launched, int* peer_exchange) {
    if (!c) return fail("null context");
    if (launched) *launched = c->last_matvecs_issued;
    if (peer_exchange) *peer_exchange = c->xchg_ok ? 1 : 0;
    return 0;
}

extern "C" int gcx_last_run_ms(gcx_ctx* c, double* ms) {
    *ms = c->last_run_ms;
    return 0;
}

extern "C" int gcx_launch_count(gcx_ctx* c, int64_t* launches) {
    *launches = c->launches;
    return 0;
}
