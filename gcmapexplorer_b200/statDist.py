"""Drop-in for the Markov-chain part of ``gcMapExplorer/lib/statDist.py``: transition
probability matrix and stationary distribution of a contact map, computed on a B200 through libgcxnorm.so.

* ``calculateTransitionProbabilityMatrix`` / ``transitionProbabilityMatrixForCCMap`` (lib/statDist.py:50-174): two
  elementwise passes over the resident map (minimum pass, output pass) after the row-sum pass shared with the mask.
* ``stationaryDistributionByEigenDecomp`` / ``statDistrByEigenDecompForCCMap`` (lib/statDist.py:272-362, 436-498):
  the reference runs a dense LAPACK eigendecomposition of the transposed kept block (O(n^3), float64 copy of the
  block) and keeps the first eigenvector; here the same vector -- the dominant left eigenvector of a positive
  matrix -- is found by power iteration with one read of the resident float32 matrix per step.  Agreement with
  the reference on its own outputs: <= 1e-9 relative (tests/test_gpu_parity.py, golden fixtures sd_*).

The ``*ForGCMap`` wrappers need the HDF5 container and raise NotImplementedError (h5py is not in this image); the
dense matrix-power routines (lib/statDist.py:500-822) are dead code in the reference (their only caller is commented
out, :825-839) and are not provided.
"""
import logging

import numpy as np

from . import ccmap as cmp
from . import ccmapHelpers as cmh
from . import util
from ._lib import GcxError
from .device import DeviceMap

logger = logging.getLogger('StatDist')
logger.setLevel(logging.INFO)

#: details of the most recent stationary-distribution run (iterations, last L1 change, device ms)
last_run_info = {}


def calculateTransitionProbabilityMatrix(A, Out=None, bNoData=None):
    """Row-stochastic transition matrix of `A` (lib/statDist.py:50-99): kept rows divided by their sum; every entry of a
    no-data row or column and every zero becomes the smallest non-zero probability.  Fills `Out` and returns None, or
    returns a new array of A's dtype.  The arithmetic is float32 (the dtype of every CCMAP matrix, lib/ccmap.py:42);
    like the reference (:90) a missing `bNoData` is a TypeError."""
    if bNoData is None:
        raise TypeError("'NoneType' object is not subscriptable")        # what lib/statDist.py:90 does with bNoData=None
    keep = ~np.asarray(bNoData, dtype=bool)
    with DeviceMap.from_host(A) as dm:
        dm.set_mask(keep)
        dm.transition(in_place=True)
        if Out is None:
            res = dm.download(0)
            return res if A.dtype == np.float32 else res.astype(A.dtype)
        if isinstance(Out, np.ndarray) and Out.dtype == np.float32 and Out.shape == A.shape and Out.strides[1] == 4:
            dm.download(0, out=Out)
        else:
            Out[:] = dm.download(0)
    return None


def transitionProbabilityMatrixForCCMap(ccMap, outFile=None, percentile_threshold_no_data=None, threshold_data_occup=None, workDir=None):
    """Transition probability matrix of a CCMAP or .ccmap file (lib/statDist.py:101-174).  Returns a new temporary
    CCMAP (unreadable, like the reference :155) with matrix, minvalue, maxvalue and bNoData set, or None when
    `outFile` is given.  The mask is recomputed when a threshold is given or the input carries none (:138-141)."""
    ccMapObj, ccmapType = cmp.checkCCMapObjectOrFile(ccMap, workDir=workDir)
    ccMapObj.make_readable()
    with DeviceMap.from_host(ccMapObj.matrix) as dm:
        if percentile_threshold_no_data is not None or threshold_data_occup is not None or ccMapObj.bNoData is None:
            bNoData = ~cmh.device_mask(dm, percentile_threshold_no_data, threshold_data_occup)
        else:
            bNoData = np.asarray(ccMapObj.bNoData, dtype=bool)
        dm.set_mask(~bNoData)
        mn, mx = dm.transition(in_place=False)
        normCCMap = cmp.result_like(dm, 1, ccMapObj)
    normCCMap.bNoData = bNoData
    normCCMap.minvalue = float(mn)
    normCCMap.maxvalue = float(mx)
    normCCMap.make_unreadable()
    if ccmapType == 'File':
        del ccMapObj
    if outFile is not None:
        cmp.save_ccmap(normCCMap, outFile, compress=True)
        del normCCMap
        return None
    return normCCMap


def stationaryDistributionByEigenDecomp(prob_matrix, bNoData, tol=1e-13, max_iter=100000):
    """Stationary distribution of a transition matrix (lib/statDist.py:436-498): float64[n], no-data bins hold the
    smallest component, the whole vector sums to 1.  `tol` / `max_iter` steer the power iteration that replaces the
    reference's dense eigendecomposition (L1 change of the probability vector / cap on matrix products)."""
    bNoData = np.asarray(bNoData, dtype=bool)
    keep = ~bNoData
    with DeviceMap.from_host(prob_matrix) as dm:
        dm.set_mask(keep)
        r = dm.stationary(tol, max_iter)
    last_run_info.clear()
    last_run_info.update(n=int(keep.size), kept=int(keep.sum()), iterations=r['iterations'], delta=r['delta'], device_ms=r['ms'])
    distrib = r['x'][keep]
    prob_eigenvector = np.zeros(prob_matrix.shape[0])
    prob_eigenvector[bNoData] = np.amin(distrib)                       # :486-494
    prob_eigenvector[keep] = distrib
    return prob_eigenvector / prob_eigenvector.sum()                   # :496


def statDistrByEigenDecompForCCMap(ccMap, chrom=None, hdf5Handle=None, compression='lzf', workDir=None):
    """Stationary distribution of a transition-matrix CCMAP or file (lib/statDist.py:272-362): outliers (modified
    z-score > 8 among the bins with data) are zeroed and the vector renormalised.  With `hdf5Handle` (any object with
    the reference handler's open / addDataByArray / close) the vector and its coarse versions are stored there and
    None is returned; numeric failures are logged and give None (:313-316)."""
    if hdf5Handle is not None and chrom is None:
        logger.info('No chromosome name provided. Skipping calculation...')
        return None
    ccMapObj, ccmapType = cmp.checkCCMapObjectOrFile(ccMap, workDir=workDir)
    ccMapObj.make_readable()
    try:
        sdist = stationaryDistributionByEigenDecomp(ccMapObj.matrix, ccMapObj.bNoData)
    except (KeyboardInterrupt, SystemExit, GcxError):
        raise
    except Exception:
        logger.warning('Not able to perform eigendecomposition... Aborting for {0}...'.format(chrom))
        ccMapObj.make_unreadable()
        return None
    ccMapObj.make_unreadable()

    minvalue = np.amin(sdist)                                           # :320-326: the no-data fill is masked,
    filtered = util.detectOutliersMasked1D(np.ma.masked_equal(sdist, minvalue), thresh=8)
    sdist = filtered.filled(0.0)[:]
    sdist = sdist / np.sum(sdist)                                       # :329

    resolution = util.binsizeToResolution(ccMapObj.binsize)
    if ccmapType == 'File':
        del ccMapObj
    if hdf5Handle is None:
        return sdist
    hdf5Handle.open()                                                   # :338-358
    for key in ('maximum', 'average', 'sum'):
        hdf5Handle.addDataByArray(chrom, resolution, key, sdist, compression=compression)
    for level in [2, 4, 5, 8, 10, 16, 20, 32, 64]:
        coarse = {'maximum': cmp.downSample1D(sdist, level=level, func='max'),
                  'average': cmp.downSample1D(sdist, level=level, func='mean'),
                  'sum': cmp.downSample1D(sdist, level=level, func='sum')}
        new_resolution = util.binsizeToResolution(util.resolutionToBinsize(resolution) * level)
        for key in ('maximum', 'average', 'sum'):
            hdf5Handle.addDataByArray(chrom, new_resolution, key, coarse[key], compression=compression)
        if coarse['sum'].shape[0] < 250:
            break
    hdf5Handle.close()
    return None


def _gcmap_unavailable(name):
    def f(*args, **kwargs):
        raise NotImplementedError(
            name + ": the HDF5 .gcmap container (gcMapExplorer/lib/gcmap.py) needs h5py, which is not available in this "
            "image; call the *ForCCMap function per map (SURVEY.md section 8f, rows N1 / N3).")
    f.__name__ = name
    return f


transitionProbabilityMatrixForGCMap = _gcmap_unavailable('transitionProbabilityMatrixForGCMap')
statDistrByEigenDecompForGCMap = _gcmap_unavailable('statDistrByEigenDecompForGCMap')
