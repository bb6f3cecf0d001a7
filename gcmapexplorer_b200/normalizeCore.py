"""Mirror of ``gcMapExplorer/lib/normalizeCore.pyx`` -- same function names, arguments and in-place /
``Out=`` conventions; the arithmetic runs in sm_100a kernels through libgcxnorm.so (no CPU fallback)."""
import numpy as np

from .device import DeviceMap

#: filled by the last performIterativeCorrection call: passes, matvecs, l1, bias, device ms
last_ic_info = {}


def _write_back(dm, which, target):
    """Download into `target` (ndarray / r+ memmap); falls back to a bounce array for odd layouts."""
    if isinstance(target, np.ndarray) and target.dtype == np.float32 and target.ndim == 2 and target.strides[1] == 4 \
            and target.strides[0] % 4 == 0 and target.flags.writeable:
        dm.download(which, out=target)
    else:
        target[:] = dm.download(which)


def performIterativeCorrection(matrix, tol, iteration):
    """normalizeCore.pyx:34-60: double-sided ICE on `matrix` **in place**; returns None.

    The recurrence is evaluated with one fp64-accumulating matvec per pass and the matrix
    is rewritten once at the end as A_ij/(B_i B_j)."""
    with DeviceMap.from_host(matrix) as dm:
        dm.set_mask(np.ones(matrix.shape[0], dtype=bool))
        r = dm.ic(tol, iteration)
        dm.apply_sym(None, masked_mode=1, in_place=True)
        _write_back(dm, 0, matrix)
    last_ic_info.clear()
    last_ic_info.update(r)


def performVCNormalization(A, Out=None, sqroot=False, bNoData=None):
    """normalizeCore.pyx:63-89 in its fp32 arithmetic.  Returns Out only when Out was None."""
    n = A.shape[0]
    keep = np.ones(n, dtype=bool) if bNoData is None else ~np.asarray(bNoData, dtype=bool)
    masked_prior = None
    if Out is not None and not keep.all():
        # rows the reference skips (:74-76) keep whatever the caller put in Out, then get the column
        # division (:79-83); the kernel assumes Out started as a copy of A (lib/normalizer.py:1117), so
        # remember the caller's rows to honour any other prior content.
        masked_prior = np.array(Out[~keep], dtype=np.float32)
    with DeviceMap.from_host(A) as dm:
        dm.set_mask(keep)
        _, _, csum = dm.vc(sqroot=sqroot, in_place=False)
        if Out is None:
            res = dm.download(1)
            res[~keep] = 0.0          # Out=None starts from zeros (:68-70): skipped rows stay 0
            return res
        _write_back(dm, 1, Out)
    if masked_prior is not None:
        with np.errstate(divide="ignore", invalid="ignore"):
            masked_prior[:, keep] = masked_prior[:, keep] / csum[keep]
            if sqroot:
                masked_prior = np.sqrt(masked_prior)
        Out[~keep] = masked_prior
    return None


def _diag_apply(A, avgContacts, Out, subtract):
    with DeviceMap.from_host(A) as dm:
        dm.diag_apply(np.asarray(avgContacts, dtype=np.float64), subtract=subtract, in_place=False)
        if Out is None:
            return dm.download(1)
        _write_back(dm, 1, Out)
    return None


def normalizeByAvgContactByDivision(A, avgContacts, Out=None):
    """normalizeCore.pyx:91-121: Out_ij = A_ij / avg[|i-j|] (A_ij where avg is 0)."""
    return _diag_apply(A, avgContacts, Out, False)


def normalizeByAvgContactBySubtraction(A, avgContacts, Out=None):
    """normalizeCore.pyx:123-170: A_ij - avg[|i-j|]; negative -> 1; zero input -> 0; avg 0 -> A_ij."""
    return _diag_apply(A, avgContacts, Out, True)


def applyNormVectors(A, c1_norm, c2_norm, Out=None):
    """The norm-vector step of the .hic importer on a dense map (lib/hic2gcmap.py:82-98): every stored count divided by
    ``c1_norm[bin_x] * c2_norm[bin_y]`` in float64, ``inf`` where that product is zero, bins without a record (zeros) untouched.
    Same calling convention as the other cores: fills `Out` when given, else returns the new array."""
    with DeviceMap.from_host(A) as dm:
        dm.apply_norm_vectors(c1_norm, c2_norm, in_place=False)
        if Out is None:
            return dm.download(1)
        if isinstance(Out, np.ndarray) and Out.dtype == np.float32 and Out.shape == tuple(A.shape):
            dm.download(1, out=Out)
        else:
            Out[:] = dm.download(1)
    return None
