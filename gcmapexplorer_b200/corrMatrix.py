"""Drop-in for ``gcMapExplorer/lib/corrMatrix.py`` + ``_corrMatrixCore``: the
pairwise-masked Pearson correlation (and covariance) matrix of a contact map on a B200.

The reference loops over all row pairs on the CPU (OpenMP, two passes per pair: lib/corrMatrixCoreSRC.c:8-58, 132-192).
Here the pair sums are three GEMM-shaped products -- N = I I^T, Sx = X0 I^T, Sxy = X0 X0^T with X0 the rows with masked
entries zeroed and I the 0/1 indicator -- and cov = (Sxy - Sx Sx^T / N) / N.  The counts run on the bf16 tensor cores (exact);
on an integer-valued map (raw counts) and on a float32-valued one (anything but log space) the other two run as exact int8
slice products, otherwise as a float64 GEMM and SYRK.
The row variances repeat the reference's sequential arithmetic, so its ``var == 0`` decisions are reproduced exactly.  Agreement with the compiled reference core: ~1e-15
absolute on the test maps (tests assert 1e-10).

Quirks kept on purpose: a row's variance uses only its own mask while the covariance uses the pair's common mask, so
coefficients can leave [-1, 1]; the diagonal of the correlation matrix is 0; pairs with at most 3 common entries give 0;
with ``logspace`` an entry equal to 1 becomes log10(1) = 0 = the default mask value and drops out.
"""
import logging

import numpy as np

from . import ccmap as cmp
from ._lib import GcxError
from .ccmap import checkCCMapObjectOrFile
from .device import DeviceMap

__all__ = ['calculateCorrMatrix', 'calculateCovMatrix', 'calculateCorrMatrixForCCMap', 'calculateCorrMatrixForGCMaps']

logger = logging.getLogger('CorrMatrix')
logger.setLevel(logging.INFO)


def _pair_matrix(in_array, out, maskvalue, covariance):
    if in_array is None:
        raise TypeError("Argument 'in_array' must not be None")
    with DeviceMap(32) as dm:                      # a context; the float64 input is staged by gcx_corr_matrix_f64 itself
        if out is None:
            return dm.corr_f64(in_array, maskvalue, covariance)
        dm.corr_f64(in_array, maskvalue, covariance, out=out)
    return None


def calculateCorrMatrix(in_array, out=None, maskvalue=None):
    """Correlation between the rows of a square float64 array (lib/_corrMatrixCore.pyx:49-98).  Entries equal to
    ``maskvalue`` (rounded to float32, :88) are left out pair by pair.  Returns the matrix, or None when `out` is given."""
    return _pair_matrix(in_array, out, maskvalue, False)


def calculateCovMatrix(in_array, out=None, maskvalue=None):
    """Covariance between the rows of a square float64 array (lib/_corrMatrixCore.pyx:100-146), diagonal included."""
    return _pair_matrix(in_array, out, maskvalue, True)


def calculateCorrMatrixForCCMap(inputCCMap, logspace=False, maskvalue=0.0, vmin=None, vmax=None, outFile=None, workDir=None):
    """Correlation matrix of a CCMAP or .ccmap file (lib/corrMatrix.py:48-159).  Entries <= vmin / >= vmax are set to
    `maskvalue` first (:96-99); the bins flagged in `bNoData` are left out and stay 0 in the result (:108, :124-133);
    minvalue / maxvalue are -1 / 1 (:135-136).  Returns a new temporary CCMAP (editable, like the reference), None when
    `outFile` is given, and None after a warning on numeric errors (:145-151) -- e.g. a map without `bNoData`."""
    ccMapObj, ccmapType = checkCCMapObjectOrFile(inputCCMap, workDir=workDir)
    try:
        ccMapObj.make_readable()
        bNoData = ccMapObj.bNoData
        keep = ~bNoData                                   # TypeError for a map without bNoData, as in the reference (:108)
        with DeviceMap.from_host(ccMapObj.matrix) as dm:
            dm.set_mask(keep)
            dm.corr(maskvalue, logspace, vmin, vmax)
            corrCCMap = cmp.result_like(dm, 1, ccMapObj)
        corrCCMap.bNoData = np.asarray(bNoData, dtype=bool).copy()
        corrCCMap.minvalue = -1
        corrCCMap.maxvalue = 1
    except (KeyboardInterrupt, SystemExit, GcxError):
        raise
    except Exception as e:
        logger.warning(e)
        logger.warning('Error in Correlation Matrix Calculation !!!')
        return None
    if outFile is not None:
        cmp.save_ccmap(corrCCMap, outFile, compress=True)
        del corrCCMap
        return None
    return corrCCMap


def calculateCorrMatrixForGCMaps(*args, **kwargs):
    raise NotImplementedError("calculateCorrMatrixForGCMaps: the HDF5 .gcmap container (gcMapExplorer/lib/gcmap.py) needs h5py, which is "
                              "not available in this image; call calculateCorrMatrixForCCMap per map (SURVEY.md section 8f, rows N1 / N4).")
