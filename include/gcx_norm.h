/* gcx_norm.h -- C ABI of libgcxnorm.so: B200-native (sm_100a) contact-map normalization.
 *
 * This is the drop-in boundary for the hot path of rjdkmr/gcMapExplorer's normalizer.  The reference
 * has no FFI layer: its seam is the three Cython modules lib/normalizeCore.pyx, lib/normalizeKnightRuiz.pyx,
 * lib/ccmapHelpers.pyx plus lib/cmstats.py:getAvgContactByDistance, called from lib/normalizer.py.
 * Each entry point below names the reference interface (file:line, relative to
 * /root/reference/gcMapExplorer/) it replaces.  INTEGRATION.md shows the ctypes binding.
 *
 * Conventions
 *   - plain pointers and sizes only; all pointers are HOST pointers unless stated otherwise;
 *   - every function returns 0 on success, non-zero on failure; gcx_last_error() gives the message
 *     (thread-local);
 *   - a gcx_ctx owns one CUDA device, one stream, one resident map (the rank's row block of the dense
 *     float32 n x n matrix, pitch padded to 128 B), an optional output map and all O(n) vectors;
 *   - nothing here falls back to the CPU: without a CUDA device every call fails.
 */
#ifndef GCX_NORM_H
#define GCX_NORM_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct gcx_ctx gcx_ctx;

/* ---- library / context ------------------------------------------------------------------------ */
const char* gcx_last_error(void);
int gcx_version(void);
int gcx_device_count(int* count);
int gcx_ctx_create(int device, gcx_ctx** out);
int gcx_ctx_destroy(gcx_ctx* ctx);
int gcx_sync(gcx_ctx* ctx);
/* device properties the benchmark reports: SM count, total/free HBM bytes */
int gcx_device_info(gcx_ctx* ctx, int* sm_count, int64_t* hbm_total, int64_t* hbm_free);

/* pinned host memory for callers that want full-speed PCIe copies (bench.py e2e leg) */
int gcx_host_alloc(int64_t bytes, void** ptr);
int gcx_host_free(void* ptr);

/* ---- map residency ----------------------------------------------------------------------------
 * Replaces the CCMAP memmap copies of lib/normalizer.py:308, 561, 842, 1117 (ccMapObj.copy()) and the
 * compaction matrix[bNZ,:][:,bNZ] (lib/normalizer.py:329, 589) / remove_zeros (lib/ccmapHelpers.pyx:241-301):
 * the map is uploaded once, never compacted -- masked bins are handled by zeroing vector entries.
 * n = global number of bins, [row0, row0+nloc) = the row block this context owns (single GPU: 0, n). */
int gcx_map_create(gcx_ctx* ctx, int64_t n, int64_t row0, int64_t nloc);
int gcx_map_upload(gcx_ctx* ctx, const float* host_rows, int64_t host_ld);          /* host_rows -> first local row */
int gcx_map_download(gcx_ctx* ctx, int which, float* host_rows, int64_t host_ld);   /* which: 0 input, 1 output, 2 second output */
/* Streaming residency (SURVEY.md 8f row N1 -- replaces the memmap-to-memmap copies and np.memmap page faults of lib/ccmap.py:264-300,
 * 404-567 on the way in and out):
 *   gcx_map_upload_rows    local rows [row_begin, row_begin + nrows) from a host chunk (a gzip / HDF5 reader hands the map over
 *                          piecewise); with_stats != 0 computes the chunk's row statistics (lib/ccmapHelpers.pyx:190-201) right
 *                          behind its copy, so the mask pass overlaps the upload and the map is not read again for it;
 *   gcx_map_upload_finish  after the last chunk: publishes those statistics (gcx_row_stats then returns them without a pass);
 *   gcx_map_upload_file    the local row block from a raw float32 row-major file (.npbin, lib/ccmap.py:233) at file_offset:
 *                          a pool of pread threads fills two pinned stages while the previous stage travels to the device;
 *   gcx_map_download_file  a result map into such a file (created / sized as needed), parallel pwrite behind each stage. */
int gcx_map_upload_rows(gcx_ctx* ctx, const float* host_rows, int64_t host_ld, int64_t row_begin, int64_t nrows, int with_stats);
int gcx_map_upload_finish(gcx_ctx* ctx);
int gcx_map_upload_file(gcx_ctx* ctx, const char* path, int64_t file_offset, int with_stats);
int gcx_map_download_file(gcx_ctx* ctx, int which, const char* path, int64_t file_offset);
/* Device-resident twin: dev_rows points to memory of the context's device owned by the caller (another library's
 * tensor, an importer kernel): nloc rows of n floats, row pitch dev_ld floats; one device-to-device copy, no host round trip. */
int gcx_map_adopt_device(gcx_ctx* ctx, const float* dev_rows, int64_t dev_ld);
int gcx_map_export_device(gcx_ctx* ctx, int which, float* dev_rows, int64_t dev_ld);
/* asynchronous download into PINNED host memory on the context's copy stream: returns at once, overlaps with
 * kernels enqueued afterwards (e.g. the VC result streams out while IC iterates); gcx_download_wait blocks until
 * the last async download has landed.  Kernels that later overwrite the source buffer wait for it on the device. */
int gcx_map_download_async(gcx_ctx* ctx, int which, float* pinned_host_rows, int64_t host_ld);
int gcx_download_wait(gcx_ctx* ctx);
/* device-side synthetic map: symmetric Poisson(scale*(d+1)^-alpha), empty bins, diag >= 1 */
int gcx_map_synth(gcx_ctx* ctx, uint64_t seed, double scale, double alpha, double empty_frac);
/* vmin/vmax clamp, lib/normalizer.py:299-301, 567-569, 827-829, 1105-1107: x<=vmin or x>=vmax -> 0 */
int gcx_map_clamp(gcx_ctx* ctx, int has_vmin, float vmin, int has_vmax, float vmax);
/* scaleUpInput, lib/normalizer.py:832-834: matrix *= factor (float32) */
int gcx_map_scale(gcx_ctx* ctx, float factor);
/* make the output map the new input (used to chain normalizations without a host round trip) */
int gcx_map_swap(gcx_ctx* ctx);

/* ---- mask -------------------------------------------------------------------------------------
 * Replaces the row loop of ccmapHelpers.get_nonzeros_index (lib/ccmapHelpers.pyx:190-201): per row the
 * float64 row sum and the int32 number of zero entries.  Outputs are the FULL n-vectors (allgathered when
 * sharded).  The empty-row rule (:192-195) and the percentile / occupancy comparisons (:204-237) are
 * integer work on n values and stay on the host (NumPy itself) so the mask is bit-exact. */
int gcx_row_stats(gcx_ctx* ctx, double* rowsum /* n, may be NULL */, int32_t* zero_count /* n, may be NULL */);
int gcx_set_mask(gcx_ctx* ctx, const uint8_t* keep /* n; 1 = bin takes part */);
/* flags[n]: 1 where the row holds negative entries.  lib/ccmapHelpers.pyx:192 calls a row empty when its FLOAT32 sum is 0.0; on a
 * non-negative map that is "all entries zero" = (rowsum == 0), on a signed map (o-e output, correlation map) the float32 sum of a
 * row can cancel exactly where the float64 sum does not -- the caller re-checks the flagged rows with NumPy's float32 sum. */
int gcx_row_negative_flags(gcx_ctx* ctx, int32_t* flags);

/* ---- Iterative correction ---------------------------------------------------------------------
 * Replaces normalizeCore.performIterativeCorrection(matrix, tol, iteration) (lib/normalizeCore.pyx:34-60).
 * One fused matvec per pass, fp64 accumulation; the matrix is not rewritten (see gcx_apply_sym_scale).
 * bias: n doubles (entries of masked bins are 1).  passes: number of reference passes (502 at the cap of
 * iteration=500).  matvecs = passes + 1. */
int gcx_ic_run(gcx_ctx* ctx, double tol, int iteration, double* bias, int* passes, double* last_l1, int* matvecs);

/* ---- Knight-Ruiz ------------------------------------------------------------------------------
 * Replaces KnightRuizNorm.__init__/run_for_RAM (lib/normalizeKnightRuiz.pyx:74-191, 271-286) with
 * identical control flow, device-resident.  x: n doubles (masked bins 0). */
int gcx_kr_run(gcx_ctx* ctx, double tol, double delta, double Delta, double* x, int* outer, int* mvp, double* rout);

/* ---- final apply ------------------------------------------------------------------------------
 * out_ij = f32((s_lo * A_ij) * s_hi) on kept x kept, lo/hi = min/max(i,j): GenNormMatrixRAM
 * (lib/normalizeKnightRuiz.pyx:207-227) and the accumulated effect of lib/normalizeCore.pyx:48-53.
 * s: n doubles on the host, or NULL to use the vector left resident by the last gcx_ic_run (1/bias) or
 * gcx_kr_run (x).  masked_mode: 0 -> masked rows/cols are 0 (KR, lib/normalizer.py:308),
 * 1 -> masked rows/cols keep the input value (IC write-back quirk, lib/normalizer.py:601-613).
 * minv/maxv: min non-zero / max over the kept x kept block (lib/normalizer.py:596-598, .pyx:214-216). */
int gcx_apply_sym_scale(gcx_ctx* ctx, const double* s, int masked_mode, int in_place, float* minv, float* maxv);

/* ---- Vanilla coverage -------------------------------------------------------------------------
 * Replaces normalizeCore.performVCNormalization (lib/normalizeCore.pyx:63-89) in its fp32 arithmetic.
 * Needs gcx_row_stats + gcx_set_mask first.  csum: n floats out (may be NULL). */
int gcx_vc_run(gcx_ctx* ctx, int sqroot, int in_place, float* minv, float* maxv, float* csum);

/* ---- .hic import: precomputed norm vectors (second half) ---------------------
 * Replaces the per-record arithmetic of lib/hic2gcmap.py:82-98 on a resident map: out_ij = f32(f64(A_ij) / (c1[i] * c2[j])), +inf
 * where c1[i] * c2[j] == 0, zeros (bins without a record) untouched; c1 / c2 are the KR / VC / VC_SQRT vectors a .hic file stores
 * per chromosome and resolution (lib/hicparser.py:358-412), n doubles each (c1 indexes rows, c2 columns). */
int gcx_apply_norm_vectors(gcx_ctx* ctx, const double* c1, const double* c2, int in_place, float* minv, float* maxv);

/* ---- fused tail: symmetric scale + vanilla coverage from ONE read of the map ----------------------
 * For callers that want both normalizations of a map (bench.py's step, the batch driver): the result of gcx_apply_sym_scale
 * (resident scale and mask of the last gcx_ic_run / gcx_kr_run, not in place) goes to output 1 and the result of gcx_vc_run under
 * the mask keep_vc to output 2 (gcx_map_download / gcx_map_checksum with which = 2) -- bit-identical to the two separate
 * calls, 12 n^2 bytes of HBM traffic instead of 16 n^2.  Replaces lib/normalizeKnightRuiz.pyx:207-227 / lib/normalizeCore.pyx:48-53
 * and lib/normalizeCore.pyx:63-89 together. */
int gcx_apply_sym_vc(gcx_ctx* ctx, const uint8_t* keep_vc, int masked_mode, int sqroot, float* min_s, float* max_s, float* min_v, float* max_v);

/* ---- distance-frequency (MCFS) ----------------------------------------------------------------
 * gcx_diag_stats replaces cmstats.getAvgContactByDistance (lib/cmstats.py:569-605) for one map:
 * median != 0 -> per-diagonal median of non-zero entries, else mean; avg: n doubles.
 * gcx_diag_apply replaces normalizeByAvgContactByDivision / BySubtraction (lib/normalizeCore.pyx:91-170). */
int gcx_diag_stats(gcx_ctx* ctx, int median, double* avg);
int gcx_diag_apply(gcx_ctx* ctx, const double* avg, int subtract, int in_place, float* minv, float* maxv);

/* Batched form for MANY resident maps on one device (the 24 chromosomes of a genome, BASELINE configs[3]; the loop over maps of
 * normalizeGCMapByMCFS, lib/normalizer.py:1013-1030): gcx_diag_stats + gcx_diag_apply (not in place) for every map, each stage ONE
 * grid over all maps instead of ~14 launches and two host synchronisations per map.  ctxs: `count` single-rank contexts with
 * resident maps; results as per map: output map 1, the statistic resident (gcx_diag_avg_read), minv/maxv[count].  Bit-identical
 * to the per-map calls. */
int gcx_mcfs_batch(gcx_ctx** ctxs, int count, int median, int subtract, float* minv, float* maxv);
int gcx_diag_avg_read(gcx_ctx* ctx, double* avg);
/* getAvgContactByDistance for a LIST of maps (lib/cmstats.py:572-578): per distance the non-zero entries of that diagonal of every
 * map that is large enough are pooled, then median (one radix selection fed by all maps) or mean.  avg: nmax doubles, nmax = bins of
 * the largest map. */
int gcx_diag_stats_pooled(gcx_ctx** ctxs, int count, int median, double* avg, int64_t nmax);

/* ---- compaction ---------------------------------------------------------------------------------
 * Replaces the element loop of ccmapHelpers.remove_zeros (lib/ccmapHelpers.pyx:286-299): host_out[ci][cj] =
 * (double) A[kept[ci]][kept[cj]], nk x nk float64.  Only for callers of the inner seam -- the normalizers never compact. */
int gcx_compact_f64(gcx_ctx* ctx, const int32_t* kept, int64_t nk, double* host_out);

/* ---- coarsening ---------------------------------------------------------
 * Replaces ccmap.downSample2DMap (lib/ccmap.py:737-842) on the resident map: out is (ceil((n-1)/level)+1)^2 floats on the
 * host; method 0 sum, 1 mean, 2 max; NumPy's reduction order and float32/float64 promotion are reproduced, so the result is
 * bit-identical.  minv/maxv: min non-zero / max of the output. */
int gcx_downsample(gcx_ctx* ctx, int level, int method, float* host_out, int64_t out_n, float* minv, float* maxv);

/* ---- transition probabilities and stationary distribution -----------------
 * gcx_transition_run replaces statDist.calculateTransitionProbabilityMatrix (lib/statDist.py:50-99) on the resident map
 * with the resident mask: kept rows divided by their row sum in float32, every no-data row/column entry and every zero
 * set to the smallest non-zero quotient.  Result in the output map (or in place).  minv/maxv: min non-zero / max.
 * gcx_stationary_run replaces statDist.stationaryDistributionByEigenDecomp (lib/statDist.py:436-498) up to its scatter
 * step: dominant left eigenvector of the kept block of the resident (transition) matrix, zeros of that block counted as
 * its smallest non-zero entry (:466-469), by power iteration in float64 instead of a dense eigendecomposition.
 * x: n doubles, sum 1 over the kept bins, 0 on no-data bins (the host applies the min-fill of :486-496).
 * Stops when the L1 change of x drops below tol, stalls at the rounding floor, or after max_iter products. */
int gcx_transition_run(gcx_ctx* ctx, int in_place, float* minv, float* maxv);
int gcx_stationary_run(gcx_ctx* ctx, double tol, int max_iter, double* x, int* iterations, double* last_delta);

/* ---- correlation matrix ---------------------------------------------------
 * gcx_corr_run replaces the numeric part of corrMatrix.calculateCorrMatrixForCCMap (lib/corrMatrix.py:92-135) on the resident
 * map with the resident mask: entries <= vmin / >= vmax become the mask value, the kept block is transposed and promoted to
 * float64, optionally log10'd (infinities -> 0), and the pairwise-masked Pearson correlation of its rows
 * (lib/corrMatrixCoreSRC.c:8-58, 102-192; mask value rounded to float32 like lib/_corrMatrixCore.pyx:88) is scattered into the
 * output map (no-data rows/columns and the diagonal 0, NaN -> 0).  The pair sums are three library GEMMs (cuBLAS, dlopen'ed):
 * exact counts from a bf16 indicator on the tensor cores, a float64 GEMM and a float64 SYRK.
 * gcx_corr_matrix_f64 replaces _corrMatrixCore.calculateCorrMatrix (lib/_corrMatrixCore.pyx:49-98): in/out m x m float64 on
 * the host, rows of `in` are the variables; has_mask = 0 is maskvalue=None; covariance != 0 gives calculateCovMatrix instead
 * (lib/_corrMatrixCore.pyx:100-146, diagonal included).  Works on a context without a resident map. */
int gcx_corr_run(gcx_ctx* ctx, int has_mask, double maskvalue, int logspace, int has_vmin, float vmin, int has_vmax, float vmax);
int gcx_corr_matrix_f64(gcx_ctx* ctx, const double* in, int64_t m, int has_mask, double maskvalue, int covariance, double* out);
/* Which route the last correlation run took for the value products (measurement only; the reference has no counterpart):
 * slices > 0: integer-valued map, that many exact 7-bit int8 slices; slices < 0: float32-valued map, -slices exact slices with one
 * exponent per row; 0: float64 GEMM.  sxy_sliced: 1 when Sxy came from slice products too, 0 when it came from the float64 SYRK/GEMM. */
int gcx_corr_last_route(gcx_ctx* ctx, int* slices, int* sxy_sliced);
/* The int8 products of the correlation path run on this library's own tensor-core kernel (k_i8gemm_nt in csrc/gcx_i8gemm.cuh:
 * tcgen05.mma kind::i8 with TMEM accumulators, TMA-staged 128-byte-swizzled tiles, 128 x 256 x 128 per stage; option
 * "corr_own_gemm" = 0 routes them through cublasGemmEx instead, for cross-checks).  Test / measurement hook: T = L R^T for host
 * int8 matrices mb x mb (row-major, mb a multiple of 256; NULL operands = all ones), T_host mb x mb int32 or NULL, mean device
 * milliseconds of `reps` launches. */
int gcx_debug_i8gemm(gcx_ctx* ctx, const signed char* L, const signed char* R, int mb, int32_t* T_host, int symmetric, int reps, double* ms);

/* ---- multi-GPU (row-block sharding; one context per rank) ------------------------------------
 * nccl_unique_id: the 128-byte ncclUniqueId from gcx_dist_unique_id on rank 0, broadcast by the host
 * (torch.distributed / any store).  After init, ic/kr/row_stats/diag_stats exchange their n-vectors with
 * one NCCL allgather per matrix pass on the context's stream. */
/* options (all optional):
 *   "sym"                0/1  use the upper-triangle pass when possible (default 1; env GCX_SYM)
 *   "assume_symmetric"   0/1  skip the exact symmetry test; REQUIRED for the upper-triangle pass on a sharded map, where a rank
 *                             cannot see the mirrored block (the caller vouches that the map is symmetric)
 *   "exchange_allreduce" 0/1  the row blocks are not the equal ceil(n/world) blocks: exchange n-vectors by zero-fill +
 *                             allreduce instead of allgather (set it on every rank, before gcx_map_create)
 *   "sym_blocks"         0/1  sharded upper-triangle pass: share the off-diagonal block pairs between their two ranks (default 1)
 *   "sym_balance"        0/1  ... and move the pairs' cuts until the ranks' bytes per pass are level (default 1; 0 = the plain
 *                             circulant deal; set it on every rank)
 *   "peer_exchange"      0/1  sharded upper-triangle IC pass: exchange the partial vectors by NVLink stores into peer-mapped
 *                             buffers with per-CTA flags inside the fused launch (default 1; set before gcx_map_create on every
 *                             rank; falls back to the NCCL allreduce when peer mapping is unavailable; env GCX_PEER_EXCHANGE)
 *   "ic_defer"           0/1  upper-triangle IC pass with deferred scalars (default 1; env GCX_IC_DEFER; 0 = literal order)
 *   "matvec_variant"     kernel variant for sweeps
 *   "sd_variant"         stationary distribution: 1 TMA-bulk transposed product (default), 0 per-thread loads
 *   "corr_variant"       correlation: 1 counts on the bf16 tensor cores + DSYRK + DGEMM (default), 0 three DGEMMs
 *   "corr_own_gemm"      0/1  int8 products on the library's own tcgen05 kernel (default 1) or through cublasGemmEx (0)
 *   "corr_int_path"      0/1  variant 1 on an integer-valued map (raw counts): Sx and Sxy from exact 7-bit-slice products on the
 *                             int8 tensor cores (int32 accumulation) instead of the float64 GEMM / SYRK (default 1) */
int gcx_set_option(gcx_ctx* ctx, const char* key, double value);
/* The aligned row partition (cuts at rank*n/world rounded to 1024 rows).  With it, "assume_symmetric" and
 * "exchange_allreduce", a sharded map runs the upper-triangle pass with the off-diagonal blocks dealt out in a circulant
 * pattern: rows, PCIe traffic and the symmetric work are all balanced across ranks.  Needs no GPU. */
int gcx_partition_rows(int64_t n, int world, int rank, int64_t* row0, int64_t* nloc);
int gcx_dist_unique_id(void* id128);
int gcx_dist_init(gcx_ctx* ctx, int rank, int world, const void* id128);

/* ---- measurement ------------------------------------------------------------------------------
 * gcx_profile enables CUDA-event pairs around every matrix-pass kernel on the context's stream;
 * gcx_profile_read returns launches and summed device milliseconds per kernel class since the last reset.
 * classes: 0 matvec, 1 apply_sym, 2 vc_apply, 3 row_stats, 4 diag_reduce, 5 diag_apply, 6 vector epilogue */
#define GCX_NCLASS 8
int gcx_profile(gcx_ctx* ctx, int enable);
int gcx_profile_read(gcx_ctx* ctx, int reset, int64_t launches[GCX_NCLASS], double ms[GCX_NCLASS]);
/* CUDA-event stopwatch on the context's stream (bench.py brackets its timed region with it) */
int gcx_timer_start(gcx_ctx* ctx);
int gcx_timer_stop(gcx_ctx* ctx, double* ms);
/* forget cached row statistics so the next mask / VC call re-reads the map (benchmark steps) */
int gcx_map_invalidate(gcx_ctx* ctx);
/* HBM bytes one matrix pass of the last ic/kr run read: 4*nloc*n for the dense pass, ~2*n*(n+1024) when the
 * upper-triangle pass was used (single rank, exactly symmetric map, n >= 2048; GCX_SYM=0 disables it) */
int gcx_pass_bytes(gcx_ctx* ctx, int64_t* bytes, int* symmetric);
/* diagnostics: cycles every CTA of the last upper-triangle pass spent in its main loop (needs GCX_SYM_DEBUG=1) */
int gcx_debug_sym_cycles(gcx_ctx* ctx, int64_t* cycles, int capacity, int* count);
/* diagnostics: per-CTA phase timestamps (ns) of one launch of the deferred-scalar IC pass, [grid][16] (needs GCX_SYM_DEBUG=1) */
int gcx_debug_sym_stamps(gcx_ctx* ctx, int64_t* stamps, int capacity, int* count);

/* Testing hook, no GPU needed: the panel plan of `rank` in the sharded upper-triangle pass of a map of n bins over `world` ranks
   (aligned partition of gcx_partition_rows).  nb[k] = bands of 4 rows, counted from the rank's first row, that it walks in
   column panel k (1024 columns); full[k] = 1 for an off-diagonal block (no triangle mask).  nch must be
   ceil(round_up(n, 32) / 1024).  balance != 0: with the cuts of the block pairs levelled as gcx_ic_run / gcx_kr_run do by default
   (option "sym_balance").  tests/test_abi_cpu.py checks that the plans of all ranks visit every unordered pair of bins once. */
int gcx_debug_sym_plan(int64_t n, int world, int rank, int balance, int64_t* nb, int32_t* full, int nch, int64_t* bytes);
/* device time (ms, CUDA events) of the last ic/kr run: matrix resident -> final vector */
int gcx_last_run_ms(gcx_ctx* ctx, double* ms);
/* matrix-vector products the last ic/kr run really launched: the deferred-scalar IC pass sees convergence one launch late and
 * runs passes + 2 products where the recurrence needs passes + 1 (what gcx_ic_run reports); peer_exchange: 1 when the sharded
 * symmetric IC pass exchanges its partial vectors through peer-mapped NVLink stores instead of an NCCL allreduce */
int gcx_last_run_products(gcx_ctx* ctx, int* launched, int* peer_exchange);
/* total kernel launches issued by this context since creation */
int gcx_launch_count(gcx_ctx* ctx, int64_t* launches);
/* micro-benchmark: `iters` back-to-back matvec launches (variant: 0 = default), device ms total */
int gcx_bench_matvec(gcx_ctx* ctx, int variant, int iters, double* ms_total);
/* 64-bit checksum of the selected map (order-independent sum of hashed entries; sharded parity checks) */
int gcx_map_checksum(gcx_ctx* ctx, int which, uint64_t* sum);

#ifdef __cplusplus
}
#endif
#endif /* GCX_NORM_H */
