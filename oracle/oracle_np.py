"""TEST INFRASTRUCTURE ONLY -- CPU (NumPy) restatement of gcMapExplorer's normalization hot path.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module; the product (gcmapexplorer_b200) never does and has no CPU fallback.

Every function cites the reference lines (relative to /root/reference/gcMapExplorer/) it restates.

PINNING.  The reference ships no tests or golden vectors for this path, so the
restatement is pinned against *outputs of the reference itself run in the build container*:
tests/golden/*.npz are produced by tests/golden/gen_golden.py from the reference's own public API
(lib/normalizer.py driven through oracle/ref_loader.py) and tests/test_oracle.py checks every
function below against them; where oracle/_ref is present the compiled reference cores are also
called live.

Arithmetic notes that matter for parity (verified against the reference under NumPy 2.3):
  * IC as-is is fp32 in place, only the (discarded) bias vector is fp64  -> ic_asis()
    The tight-tolerance target is the same recurrence evaluated in fp64   -> ic_fp64()
  * KR is fp64 arithmetic on fp32 data (np.dot upcasts)                   -> kr()
  * VC is pure fp32: row sums fp32, two fp32 divisions per element        -> vc()
  * MCFS: per-diagonal mean/median in fp64; A/avg evaluated in fp64 then stored fp32 (NumPy >= 2
    promotion of float32-array / float64-scalar)                          -> avg_contact_by_distance(), mcfs_*()
"""
import math

import numpy as np


# --------------------------------------------------------------------------------------------------
# mask  (lib/ccmapHelpers.pyx:156-239)
# --------------------------------------------------------------------------------------------------
def row_stats(matrix):
    """Per-row (is_empty, zero_count) exactly as the loop at lib/ccmapHelpers.pyx:190-201 builds them.

    A row is empty when np.sum(row) == 0.0 (:192); empty rows record a zero count of 0 (:195), every
    other row records (row == 0).sum() (:201)."""
    n = matrix.shape[0]
    empty = np.zeros(n, dtype=bool)
    zero_count = np.zeros(n, dtype=np.int64)
    for i in range(n):
        row = matrix[i]
        if np.sum(row) == 0.0:
            empty[i] = True
        else:
            zero_count[i] = int((row == 0).sum())
    return empty, zero_count


def mask_from_stats(empty, zero_count, threshold_percentile=None, threshold_data_occup=None):
    """The filter logic of lib/ccmapHelpers.pyx:204-237 applied to precomputed row statistics.

    Returns bData (True = row/column has data)."""
    n = zero_count.shape[0]
    bData = ~np.asarray(empty, dtype=bool)
    zero_count = np.asarray(zero_count)
    if threshold_percentile is not None and threshold_data_occup is not None:      # :206-207
        raise AssertionError("Both 'threshold_percentile' and 'threshold_count_ratio' cannot be used simultaneously!")
    if threshold_percentile is not None and zero_count.any():                      # :210-219
        p = np.percentile(zero_count[zero_count != 0], threshold_percentile)
        comp = int((zero_count == 0).sum())
        bData &= ~((zero_count - comp) >= p)
    if threshold_data_occup is not None and zero_count.any():                      # :222-227
        bData &= ~((zero_count / n) >= (1.0 - threshold_data_occup))
    if bData.sum() == 0:                                                           # :231-237
        if threshold_percentile is not None:
            raise ValueError("All rows contain zero values!!! Try increasing 'threshold_percentile' value or do not use it.")
        if threshold_data_occup is not None:
            raise ValueError("All rows contain zero values!!! Try decreasing 'threshold_data_occup' value or do not use it.")
    return bData


def get_nonzeros_index(matrix, threshold_percentile=None, threshold_data_occup=None):
    """lib/ccmapHelpers.pyx:156-239 (filterByDiagonal is never passed by the normalizer)."""
    empty, zc = row_stats(matrix)
    return mask_from_stats(empty, zc, threshold_percentile, threshold_data_occup)


# --------------------------------------------------------------------------------------------------
# IC  (lib/normalizeCore.pyx:34-60)
# --------------------------------------------------------------------------------------------------
def _ic_recurrence(M, tol, iteration, vec_dtype):
    """The while-loop of lib/normalizeCore.pyx:41-60 on a working matrix M (mutated in place).

    vec_dtype is the dtype numpy gives dB in the reference: float32 for the as-is run (M is float32),
    float64 for the tight oracle.  B is float64 in both (:37).  Returns (B, passes, l1_history)."""
    n = M.shape[0]
    cc = np.sum(M)                                                  # :36  contact_count
    B = np.ones((n, 1), dtype=np.float64)                           # :37
    prev_B = None
    count = 0
    passes = 0
    l1 = []
    while True:
        dB = np.sum(M, axis=0).reshape((n, 1)).astype(vec_dtype, copy=False)   # :42-43
        dB = dB / np.mean(dB[dB != 0])                              # :44
        dB[dB == 0] = 1                                             # :45
        M /= dB                                                     # :48
        M /= dB.T                                                   # :49
        B = B * dB                                                  # :51
        s = np.sum(M)
        B *= np.sqrt(s / cc)                                        # :52
        M *= cc / np.sum(M)                                         # :53
        passes += 1
        if prev_B is not None:                                      # :55-57
            d = np.abs(prev_B - B).sum()
            l1.append(float(d))
            if d < tol or count > iteration:
                break
        count += 1                                                  # :59
        prev_B = B.copy()                                           # :60
    return B[:, 0], passes, l1


def ic_asis(matrix, tol=1e-4, iteration=500):
    """As-is restatement: fp32 in-place arithmetic exactly like the reference. Returns (M32, B, passes)."""
    M = np.array(matrix, dtype=np.float32, copy=True)
    B, passes, _ = _ic_recurrence(M, tol, iteration, np.float32)
    return M, B, passes


def ic_fp64(matrix, tol=1e-4, iteration=500):
    """Tight oracle: the identical recurrence evaluated in fp64.

    Returns (M64, B, passes, l1_history); the normalized map is M64 = A_ij / (B_i * B_j)."""
    M = np.array(matrix, dtype=np.float64, copy=True)
    B, passes, l1 = _ic_recurrence(M, tol, iteration, np.float64)
    return M, B, passes, l1


def ic_fp64_matvec(matrix, tol=1e-4, iteration=500, matvec=None):
    """The recurrence of lib/normalizeCore.pyx:41-60 in fp64 WITHOUT rewriting the matrix, for sizes where ic_fp64's five
    sweeps of an n x n float64 array per pass take minutes (tests/golden/gen_golden_large.py: n = 24,926).

    The working matrix of pass k is M = A o (w w') (every step of :48-53 scales rows/columns by a vector or everything by a
    scalar), so with t = A' w (the column sums of :42 are w o t):
        dB  = w o t / mean(w o t != 0), zeros -> 1      (:42-45)
        w1  = w / dB                                    (:48-49)
        s   = w1' A w1                                  (np.sum(M) of :52-53)
        B  *= dB * sqrt(s / cc)                         (:51-52)
        w   = w1 * sqrt(cc / s)                         (:53)
    and the next pass's column sums come from the same product t1 = A' w1 that gave s (w o (A' w) = (cc/s) w1 o t1): one product
    per pass.  Checked against ic_fp64 (the literal in-place form) in tests/test_oracle.py: equal pass counts, bias to 1e-12.
    matvec(w) must return A' w in float64 (default: float64 copy of `matrix`, BLAS dgemv).
    Returns (B, passes, l1_history)."""
    n = matrix.shape[0]
    if matvec is None:
        A64 = np.asarray(matrix, dtype=np.float64)
        matvec = lambda w: A64.T.dot(w)                                # noqa: E731
    w = np.ones(n)
    t = matvec(w)
    cc = float(t.sum())                                                # :36
    colsum = w * t
    B = np.ones(n)
    prev_B = None
    count = passes = 0
    l1 = []
    while True:
        dB = colsum / np.mean(colsum[colsum != 0])                     # :44
        dB[dB == 0] = 1                                                # :45
        w1 = w / dB                                                    # :48-49
        t1 = matvec(w1)
        s = float(w1 @ t1)
        B = B * dB * math.sqrt(s / cc)                                 # :51-52
        g = cc / s                                                     # :53
        w = w1 * math.sqrt(g)
        colsum = g * (w1 * t1)
        passes += 1
        if prev_B is not None:                                         # :55-57
            d = float(np.abs(prev_B - B).sum())
            l1.append(d)
            if d < tol or count > iteration:
                break
        count += 1
        prev_B = B.copy()
    return B, passes, l1


# --------------------------------------------------------------------------------------------------
# KR  (lib/normalizeKnightRuiz.pyx:74-288, RAM variant)
# --------------------------------------------------------------------------------------------------
def kr(A, tol=1e-12, delta=0.1, Delta=3, max_outer=100000):
    """Knight-Ruiz balancing with the reference's exact control flow.

    init (:106-124), inner CG step (:126-165), outer Newton step (:167-191), driver loops (:271-286).
    A is the compacted float32 matrix; all vector arithmetic is float64 (np.dot upcasts A).
    Returns dict(x, outer, MVP, rout)."""
    n = A.shape[0]
    e = np.ones(n)
    g, etamax = 0.9, 0.1                                            # :113-114
    eta = etamax
    stop_tol = tol * 0.5                                            # :116
    x = e.copy()
    rt = tol ** 2                                                   # :119
    v = x * A.dot(x)                                                # :120
    rk = 1.0 - v
    rho_km1 = float(rk @ rk)                                        # :122
    rho_km2 = None
    rout = rho_km1
    rold = rout
    MVP = 0
    outer = 0
    Z = p = None
    while rout > rt and outer < max_outer:                          # :275
        outer += 1
        k = 0
        y = e.copy()
        innertol = max(eta ** 2 * rout, rt)                         # :279
        while rho_km1 > innertol:                                   # :281 (stale rho on entry - literal)
            k += 1
            if k == 1:                                              # :129-132
                Z = rk / v
                p = Z
                rho_km1 = float(rk @ Z)
            else:                                                   # :133-135
                beta = rho_km1 / rho_km2
                p = Z + beta * p
            w = x * A.dot(x * p) + v * p                            # :139
            alpha = rho_km1 / float(p @ w)                          # :140
            ap = alpha * p
            ynew = y + ap                                           # :144
            if np.amin(ynew) <= delta:                              # :147-153
                if delta == 0:
                    break
                ind = ap < 0
                gamma = np.amin((delta - y[ind]) / ap[ind])
                y = y + gamma * ap
                break
            if np.amax(ynew) >= Delta:                              # :154-158
                ind = ynew > Delta
                gamma = np.amin((Delta - y[ind]) / ap[ind])
                y = y + gamma * ap
                break
            y = ynew                                                # :159
            rk = rk - alpha * w                                     # :160
            rho_km2 = rho_km1
            Z = rk / v
            rho_km1 = float(rk @ Z)                                 # :163
        x = x * y                                                   # :168
        v = x * A.dot(x)                                            # :169
        rk = 1.0 - v
        rho_km1 = float(rk @ rk)
        rout = rho_km1
        MVP += k + 1                                                # :173
        rat = rout / rold                                           # :176-187
        rold = rout
        res_norm = math.sqrt(rout)
        eta_o = eta
        eta = g * rat
        if g * eta_o ** 2 > 0.1:
            eta = max(eta, g * eta_o ** 2)
        eta = max(min(eta, etamax), stop_tol / res_norm)
    return dict(x=x, outer=outer, MVP=MVP, rout=rout)


def kr_matrix(A, x):
    """GenNormMatrixRAM arithmetic (lib/normalizeKnightRuiz.pyx:207-212): D_ij = f32((x_j*A_ij)*x_i)."""
    D = (x[None, :] * A.astype(np.float64)) * x[:, None]
    return D.astype(np.float32)


def scatter_symmetric(D, bNoData, out):
    """The write-back loop of lib/normalizeKnightRuiz.pyx:218-227 / lib/normalizer.py:601-609.

    Row i of D is written to out[i, kept] *and* out[kept, i] for increasing i, so the value that
    survives at (a, b) is D[max(a,b), min(a,b)] (the lower triangle wins)."""
    kept = np.nonzero(~bNoData)[0]
    L = np.tril(D)
    S = L + np.tril(D, -1).T
    out[np.ix_(kept, kept)] = S
    return out


# --------------------------------------------------------------------------------------------------
# VC  (lib/normalizeCore.pyx:63-89)
# --------------------------------------------------------------------------------------------------
def vc(A, bNoData=None, sqroot=False, out=None):
    """Vanilla coverage in the reference's fp32 arithmetic.  `out` starts as a copy of the input
    (lib/normalizer.py:1117) so masked rows/cols keep (half-)raw values."""
    A = np.asarray(A, dtype=np.float32)
    n = A.shape[0]
    if out is None:
        out = A.copy()
    keep = np.ones(n, dtype=bool) if bNoData is None else ~np.asarray(bNoData, dtype=bool)
    csum = np.sum(A, axis=1)                                        # :72 (fp32)
    with np.errstate(divide="ignore", invalid="ignore"):
        out[keep, :] = A[keep, :] / csum[keep, None]                # :73-77
        out[:, keep] = out[:, keep] / csum[None, keep]              # :79-83
        if sqroot:
            out[:] = np.sqrt(out)                                   # :85-86
    return out, csum


# --------------------------------------------------------------------------------------------------
# MCFS  (lib/cmstats.py:522-607, lib/normalizeCore.pyx:91-170)
# --------------------------------------------------------------------------------------------------
def avg_contact_by_distance(A, stats="median"):
    """lib/cmstats.py:569-605 for one map: stat of the non-zero entries of each diagonal, fp64, 0.0
    when a diagonal has no non-zero entry (:602-603)."""
    n = A.shape[0]
    avg = np.zeros(n, dtype=np.float64)
    for d in range(n):
        diag = np.diagonal(A, offset=-d)                            # rows[d:], cols[:-d] (util.py:95-96)
        data = diag[diag != 0].astype(np.float64)                   # :577-578 (hstack with f64 empty promotes)
        if data.shape[0] >= 1:
            avg[d] = np.median(data) if stats == "median" else np.mean(data)
    return avg


def _dist_index(n):
    idx = np.arange(n)
    return np.abs(idx[:, None] - idx[None, :])


def mcfs_divide(A, avg):
    """lib/normalizeCore.pyx:91-121: out = A/avg[|i-j|] where avg != 0 else A (fp64 divide, fp32 store)."""
    A = np.asarray(A, dtype=np.float32)
    a = avg[_dist_index(A.shape[0])]
    with np.errstate(divide="ignore", invalid="ignore"):
        out = np.where(a != 0, A.astype(np.float64) / np.where(a != 0, a, 1.0), A.astype(np.float64))
    return out.astype(np.float32)


def mcfs_subtract(A, avg):
    """lib/normalizeCore.pyx:123-170: v = A-avg; v<0 -> 1; A==0 -> 0; avg==0 -> A."""
    A = np.asarray(A, dtype=np.float32)
    a = avg[_dist_index(A.shape[0])]
    v = A.astype(np.float64) - a
    v[v < 0] = 1.0                                                  # :138-140
    v[A == 0] = 0.0                                                 # :141-143
    out = np.where(a != 0, v, A.astype(np.float64))
    return out.astype(np.float32)


# --------------------------------------------------------------------------------------------------
# orchestration semantics  (lib/normalizer.py) on plain arrays
# --------------------------------------------------------------------------------------------------
def clamp(A, vmin=None, vmax=None):
    """lib/normalizer.py:299-301, 567-569, 827-829, 1105-1107: values <= vmin or >= vmax become 0."""
    A = np.array(A, dtype=np.float32, copy=True)
    if vmin is not None:
        A[A <= vmin] = 0.0
    if vmax is not None:
        A[A >= vmax] = 0.0
    return A


def _minmax_nonzero(M):
    nz = M[M != 0]
    return (float(nz.min()) if nz.size else float("nan")), float(M.max())


def normalize_ic(A, tol=1e-4, iteration=500, vmin=None, vmax=None,
                 percentile_threshold_no_data=None, threshold_data_occup=None, fp64=True):
    """lib/normalizer.py:505-647 on arrays.  fp64=True uses the tight oracle for the core.

    Returns dict(matrix f32, bNoData, minvalue, maxvalue, bias (kept bins), passes, keep)."""
    T = clamp(A, vmin, vmax)                                        # :561-570
    n = T.shape[0]
    flt = not (percentile_threshold_no_data is None and threshold_data_occup is None)   # :577-580
    if flt:
        keep = get_nonzeros_index(T, percentile_threshold_no_data, threshold_data_occup)  # :588
        sub = T[keep, :][:, keep]                                   # :589
    else:
        keep = np.ones(n, dtype=bool)
        sub = T
    if fp64:
        M, B, passes, _ = ic_fp64(sub, tol, iteration)
    else:
        M, B, passes = ic_asis(sub, tol, iteration)
    M32 = M.astype(np.float32)
    mn, mx = _minmax_nonzero(M32)                                   # :596-598 (on the compacted result)
    out = T.copy()
    if flt:
        scatter_symmetric(M32, ~keep, out)                          # :601-609; masked rows keep raw values (:610-613 is a no-op)
        bNoData = ~keep
    else:
        out = M32
        bNoData = np.all(out == 0.0, axis=0)                        # :594, :599
    return dict(matrix=out, bNoData=bNoData, minvalue=mn, maxvalue=mx, bias=B, passes=passes, keep=keep)


def normalize_kr(A, tol=1e-12, vmin=None, vmax=None,
                 percentile_threshold_no_data=None, threshold_data_occup=None):
    """lib/normalizer.py:228-395 (memory='RAM') on arrays."""
    T = clamp(A, vmin, vmax)
    keep = get_nonzeros_index(T, percentile_threshold_no_data, threshold_data_occup)   # :327
    sub = T[keep, :][:, keep]                                       # :329
    r = kr(sub, tol=tol, delta=0.1, Delta=3)                        # :319-321, :335-336
    D = kr_matrix(sub, r["x"])
    mn, mx = _minmax_nonzero(D)                                     # normalizeKnightRuiz.pyx:214-216
    out = np.zeros_like(T)                                          # :308 copy(fill=0.0)
    scatter_symmetric(D, ~keep, out)
    return dict(matrix=out, bNoData=~keep, minvalue=mn, maxvalue=mx, x=r["x"], outer=r["outer"], MVP=r["MVP"], rout=r["rout"])


def normalize_vc(A, sqroot=False, vmin=None, vmax=None,
                 percentile_threshold_no_data=None, threshold_data_occup=None):
    """lib/normalizer.py:1041-1168 on arrays."""
    T = clamp(A, vmin, vmax)
    keep = get_nonzeros_index(T, percentile_threshold_no_data, threshold_data_occup)   # :1123
    out, csum = vc(T, bNoData=~keep, sqroot=sqroot, out=T.copy())   # :1117, :1128
    mn, _ = _minmax_nonzero(out)                                    # :1131-1133
    return dict(matrix=out, bNoData=~keep, minvalue=mn, maxvalue=float(np.amax(out)), csum=csum)


def locate_significant_digit_after_decimal(value):
    """lib/util.py:53-73."""
    location, base = 0, 0
    while base <= 0:
        base = int(abs(value) * 10 ** location)
        location += 1
    return location - 1


def normalize_mcfs(A, stats="median", stype="o/e", vmin=None, vmax=None, scaleUpInput=False, minvalue_attr=None,
                   percentile_threshold_no_data=None, threshold_data_occup=None):
    """lib/normalizer.py:742-902 on arrays.  minvalue_attr is the input CCMAP's .minvalue (used by
    scaleUpInput, :832-834)."""
    if stype not in ("o/e", "o-e"):                                 # :810-812
        return None
    T = clamp(A, vmin, vmax)
    if scaleUpInput:
        to_scale = locate_significant_digit_after_decimal(minvalue_attr) + 1
        T = (T * (10 ** to_scale)).astype(np.float32)
    keep = get_nonzeros_index(T, percentile_threshold_no_data, threshold_data_occup)   # :854
    avg = avg_contact_by_distance(T, stats)                         # :858
    out = mcfs_divide(T, avg) if stype == "o/e" else mcfs_subtract(T, avg)             # :860-864
    mn, _ = _minmax_nonzero(out)                                    # :866-868
    return dict(matrix=out, bNoData=~keep, minvalue=mn, maxvalue=float(np.amax(out)), avg=avg)


def hic_norm_apply(A, c1, c2):
    """lib/hic2gcmap.py:82-98 on a dense map: a stored count (non-zero entry) becomes count / (c1[bin_x] * c2[bin_y]) in Python
    float64 arithmetic, inf where the product is 0.0 (:89-93); the matrix is float32.  Zeros are bins without a record."""
    out = np.array(A, dtype=np.float32, copy=True)
    n = A.shape[0]
    for i in range(n):
        for j in range(n):
            count = float(A[i, j])
            if count != 0.0:
                norm = float(c1[i]) * float(c2[j])
                out[i, j] = np.float32(count / norm) if norm != 0.0 else np.float32(np.inf)
    return out


# --------------------------------------------------------------------------------------------------
# synthetic maps for CPU-side tests (SURVEY.md section 8d definition; host RNG, small n only)
# --------------------------------------------------------------------------------------------------
def synth_map(n, seed=0, scale=200.0, alpha=1.0, empty_frac=0.01):
    """Symmetric Poisson(scale*(d+1)^-alpha) counts, round(empty_frac*n) (>=2) empty bins,
    diagonal >= 1 on kept bins.  float32."""
    rng = np.random.default_rng(seed)
    d = _dist_index(n).astype(np.float64)
    lam = scale * (d + 1.0) ** (-alpha)
    U = np.triu(rng.poisson(lam).astype(np.float32))
    A = U + np.triu(U, 1).T
    k = max(2, int(round(empty_frac * n))) if empty_frac > 0 else 0
    empty = rng.choice(n, size=k, replace=False) if k else np.array([], dtype=int)
    di = np.arange(n)
    A[di, di] = np.maximum(A[di, di], 1.0)
    A[empty, :] = 0.0
    A[:, empty] = 0.0
    return np.ascontiguousarray(A, dtype=np.float32)


# --------------------------------------------------------------------------------------------------
# coarsening  lib/ccmap.py:655-842
# --------------------------------------------------------------------------------------------------
def downsample_shape(shape, level=2):
    """lib/ccmap.py:712-735."""
    return (int(np.ceil(float(shape[0] - 1) / level)) + 1, int(np.ceil(float(shape[1] - 1) / level)) + 1)


def downsample_2d(M, level=2, method="sum"):
    """lib/ccmap.py:737-842 (downSample2DMap with an output matrix given): output row/column 0 stay zero; output
    (c, 1+g) reduces input rows 1+(c-1)*level ... and columns 1+g*level ...; when the column count needs padding the
    reference hstacks float64 zeros (:789-797), which silently promotes that case to float64 arithmetic."""
    n0, n1 = M.shape
    oshape = downsample_shape(M.shape, level)
    out = np.zeros(oshape, dtype=M.dtype)
    pad_inner = int((oshape[1] - 1) * level - (n1 - 1))
    red = {"mean": np.mean, "max": np.max}.get(method, np.sum)
    count, i = 1, 1
    while i < n0:
        blk = M[i:i + level, 1:]
        if pad_inner != 0:
            blk = np.hstack((blk, np.zeros((blk.shape[0], pad_inner))))       # float64 zeros -> float64 block
            group = level
        else:
            group = level if i + level < n0 else n0 - i                        # :818-826 ("remain")
        r = red(red(blk, axis=0).reshape(-1, group), axis=1)
        out[count, 1:] = r
        i += level
        count += 1
    return out


def downsample_nodata(bNoData, out_n, level):
    """lib/ccmap.py:694-702: a coarse bin has no data if any of its fine bins has none; bin 0 is never flagged."""
    b = np.zeros(out_n, dtype=bool)
    pad = int((out_n - 1) * level - (bNoData.shape[0] - 1))
    x = np.hstack((bNoData[1:], np.zeros(pad))) if pad != 0 else bNoData[1:]
    b[1:] = x.reshape(-1, level).sum(axis=1)
    return b


# --------------------------------------------------------------------------------------------------
# transition probability matrix + stationary distribution  lib/statDist.py
# --------------------------------------------------------------------------------------------------
def transition_matrix(A, bNoData, out=None):
    """lib/statDist.py:50-99 (calculateTransitionProbabilityMatrix).  Kept rows are divided by their own sum in the
    matrix dtype (:83-84); `minvalue` is the smallest non-zero quotient (:86-87); every entry of a no-data row or
    column, and every zero of a kept row, becomes `minvalue` (:89-95).  Returns (Out, minvalue)."""
    A = np.asarray(A)
    bNoData = np.asarray(bNoData, dtype=bool)
    Out = np.zeros(A.shape, dtype=A.dtype) if out is None else out
    for m in np.nonzero(~bNoData)[0]:
        sumr = A[m].sum()
        Out[m] = (A[m] / sumr)[:]
    minvalue = np.ma.masked_equal(Out, 0.0, copy=False).min()
    Out[bNoData, :] = minvalue
    Out[:, bNoData] = minvalue
    Out[Out == 0] = minvalue
    return Out, minvalue


def stationary_eig(prob_matrix, bNoData):
    """lib/statDist.py:436-498 (stationaryDistributionByEigenDecomp): the kept block in float64, zeros replaced by
    the smallest non-zero entry (:466-469), first eigenvector LAPACK returns for the transposed block (:480-481),
    scaled to sum 1, no-data bins set to the smallest component (:486-494), scaled to sum 1 again (:496)."""
    from scipy import linalg as sp_linalg
    bNoData = np.asarray(bNoData, dtype=bool)
    n = prob_matrix.shape[0]
    A = np.array((prob_matrix[~bNoData, :])[:, ~bNoData], dtype=np.float64)
    A[A == 0] = np.ma.masked_equal(A, 0.0, copy=False).min()
    _, eigvector = sp_linalg.eig(A.T)
    distrib = eigvector[:, 0]
    distrib = np.real(distrib / distrib.sum())
    full = np.zeros(n)
    full[bNoData] = np.amin(distrib)
    full[~bNoData] = distrib
    return full / full.sum()


def stationary_power(prob_matrix, bNoData, tol=1e-13, max_iter=100000):
    """The stationary distribution as the GPU computes it: power iteration x <- x P' / |x P'|_1 on
    the kept block in float64 (zeros replaced like :466-469), stopped when the L1 change drops below `tol` or stalls at
    the rounding floor.  NOT the reference's algorithm (that is `stationary_eig`); used to check iteration counts and
    at sizes where a dense eigendecomposition takes minutes.  Returns (vector like stationary_eig, iterations)."""
    bNoData = np.asarray(bNoData, dtype=bool)
    n = prob_matrix.shape[0]
    A = np.array((prob_matrix[~bNoData, :])[:, ~bNoData], dtype=np.float64)
    A[A == 0] = np.ma.masked_equal(A, 0.0, copy=False).min()
    nk = A.shape[0]
    x = np.full(nk, 1.0 / nk)
    prev = np.inf
    it = 0
    while it < max_iter:
        y = x @ A
        y /= y.sum()
        delta = np.abs(y - x).sum()
        x = y
        it += 1
        if delta < tol or (delta >= prev and delta < 1e-11):
            break
        prev = delta
    full = np.zeros(n)
    full[bNoData] = np.amin(x)
    full[~bNoData] = x
    return full / full.sum(), it


def detect_outliers_1d(points, thresh=3.5):
    """lib/util.py:100-138: modified z-score on the median absolute deviation."""
    p = points[:, None] if points.ndim == 1 else points
    median = np.median(p, axis=0)
    diff = np.sqrt(np.sum((p - median) ** 2, axis=-1))
    mad = np.median(diff)
    return 0.6745 * diff / mad > thresh


def stationary_postprocess(sdist, thresh=8):
    """lib/statDist.py:320-333: components equal to the minimum (the no-data fill) are masked, outliers of the rest
    are zeroed (lib/util.py:140-174), everything masked becomes 0, and the vector is scaled to sum 1."""
    minvalue = np.amin(sdist)
    mx = np.ma.masked_equal(sdist, minvalue)
    pts = mx.compressed().copy()
    pts[detect_outliers_1d(pts, thresh=thresh)] = 0.0
    result = np.zeros(mx.shape, dtype=mx.dtype)
    result[~np.ma.getmaskarray(mx)] = pts
    out = np.ma.masked_values(result, 0.0).filled(0.0)
    return out / np.sum(out)


# --------------------------------------------------------------------------------------------------
# correlation matrix  lib/corrMatrixCoreSRC.c, lib/_corrMatrixCore.pyx, lib/corrMatrix.py
# --------------------------------------------------------------------------------------------------
def _pair_cov(x, y, mask):
    """lib/corrMatrixCoreSRC.c:132-192: covariance over the entries where both vectors differ from the mask value
    (all entries when mask is None); 0 unless more than 3 entries are left (:171)."""
    both = np.ones(x.shape, dtype=bool) if mask is None else (x != mask) & (y != mask)
    cnt = int(both.sum())
    if cnt <= 3:
        return 0.0
    xs, ys = x[both], y[both]
    return float(np.sum((xs - xs.sum() / cnt) * (ys - ys.sum() / cnt)) / cnt)


def corr_matrix(in_array, maskvalue=None, covariance=False):
    """lib/_corrMatrixCore.pyx:49-146 around lib/corrMatrixCoreSRC.c:8-100: correlation (pairs j < i, diagonal left 0) or
    covariance (pairs j <= i) between the rows of a square float64 array; the mask value is rounded to float32 (:88).
    A row's variance uses its own mask only (:115-116), the covariance the pair's common mask."""
    X = np.asarray(in_array, dtype=np.float64)
    m = X.shape[0]
    mask = None if maskvalue is None else float(np.float32(maskvalue))
    out = np.zeros((m, m))
    var = np.array([_pair_cov(X[i], X[i], mask) for i in range(m)])
    for i in range(m):
        for j in range(i + 1 if covariance else i):
            c = _pair_cov(X[i], X[j], mask)
            if not covariance:
                c = 0.0 if (var[i] == 0 or var[j] == 0) else c / np.sqrt(var[i] * var[j])
            out[i, j] = out[j, i] = c
    return out


def corr_for_map(A, bNoData, logspace=False, maskvalue=0.0, vmin=None, vmax=None):
    """lib/corrMatrix.py:92-135: threshold to the mask value, kept block transposed in float64, optional log10 with
    infinities set to 0, correlation of the rows, NaN -> 0, scattered into a zero float32 map of the input's shape."""
    M = np.array(A, dtype=np.float32, copy=True)
    if vmin is not None:
        M[M <= vmin] = maskvalue
    if vmax is not None:
        M[M >= vmax] = maskvalue
    keep = ~np.asarray(bNoData, dtype=bool)
    X = (M[keep, :])[:, keep].T.astype(np.float64, order='C')
    if logspace:
        with np.errstate(divide='ignore', invalid='ignore'):
            X = np.log10(X)
        X[np.isinf(X)] = 0
    C = corr_matrix(X, maskvalue=maskvalue)
    C[np.isnan(C)] = 0
    out = np.zeros(M.shape, dtype=np.float32)
    out[np.ix_(keep, keep)] = C
    return out
