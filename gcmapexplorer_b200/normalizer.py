"""Drop-in for ``gcMapExplorer.lib.normalizer`` -- same function names, keyword arguments, defaults,
return values, CCMAP side effects and error conventions (reference: lib/normalizer.py), with the numeric
work done on a B200 through libgcxnorm.so.

Per call the map is streamed once into HBM, clamped / masked / iterated / rewritten there, and the result
is downloaded straight into the new CCMAP's backing file.  The reference's 2-3 extra n^2 memmap copies
(lib/normalizer.py:308, 561, 842, 1117) and the compaction ``matrix[bNZ,:][:,bNZ]`` (:329, :589) are gone:
masked bins are handled by zeroing vector entries.

Write-back quirks of the reference are reproduced on purpose:
  * IC with a filter leaves the (clamped) raw values in masked rows/columns (lib/normalizer.py:601-613);
  * VC with a filter half-normalises masked rows (lib/normalizeCore.pyx:73-83);
  * MCFS 'o-e' maps negative differences to 1 and keeps zeros at 0 (lib/normalizeCore.pyx:138-144).

What the reference never reports -- iteration / matrix-vector-product counts, final residuals, device
time -- is logged at INFO level and kept in ``last_run_info``.
"""
import logging

import numpy as np

from . import ccmap as cmp
from . import ccmapHelpers as cmh
from . import sharded
from . import util
from ._lib import GcxError
from .device import DeviceMap

logger = logging.getLogger("normalizer")
logger.setLevel(logging.INFO)

#: details of the most recent normalize* call (method, n, passes / outer / MVP, residual, device ms ...)
last_run_info = {}


def _label(ccMapObj):
    return "{0} map".format(ccMapObj.xlabel) if ccMapObj.xlabel is not None else "UNKNOWN map!!"


def _upload(ccMapObj, vmin, vmax):
    """Stream the CCMAP's float32 memmap into HBM once and apply the vmin/vmax clamp there
    (lib/normalizer.py:299-301, 567-569, 827-829, 1105-1107)."""
    ccMapObj.make_readable()
    try:
        k = sharded.gpus_for_map(int(ccMapObj.shape[0]))
        if k > 1:
            # too large for one GPU (or GCX_GPUS=k): row blocks on k GPUs of this node, one host thread each (sharded.py)
            dm = sharded.ShardedDeviceMap.from_host(ccMapObj.matrix, devices=list(range(k)))
        else:
            dm = DeviceMap.from_host(ccMapObj.matrix)
    finally:
        ccMapObj.make_unreadable()
    if vmin is not None or vmax is not None:
        dm.clamp(vmin, vmax)
    return dm


def _emit(dm, which, like, bNoData, minvalue, maxvalue):
    """New temporary CCMAP next to `like` (this package's CCMAP or the reference's, duck-typed), filled with the device result."""
    out = cmp.result_like(dm, which, like)
    out.bNoData = np.asarray(bNoData, dtype=bool)
    out.minvalue = float(minvalue)
    out.maxvalue = float(maxvalue)
    return out


def _finish(result, outFile):
    if outFile is not None:
        cmp.save_ccmap(result, outFile, compress=True)
        del result
        return None
    return result


# ---------------------------------------------------------------------------------------------------
def normalizeCCMapByKR(ccMap, memory='RAM', tol=1e-12, outFile=None, vmin=None, vmax=None,
                       percentile_threshold_no_data=None, threshold_data_occup=None, workDir=None):
    """Knight-Ruiz matrix balancing (lib/normalizer.py:228-395).  Returns the normalized CCMAP, or None when
    ``outFile`` is given.  Errors are logged and re-raised, like the reference (:390-395).
    ``memory`` accepts 'RAM' and 'HDD' (both run on the GPU); anything else raises ValueError (:353-354)."""
    ccMapObj, ccmapType = cmp.checkCCMapObjectOrFile(ccMap, workDir=workDir)
    logger.info(' KR Normalization will be done through {0}.'.format(memory))
    logger.info(' KR Normalization is in process for {0}...'.format(_label(ccMapObj)))
    delta, Delta = 0.1, 3           # :319-320
    try:
        if memory not in ('RAM', 'HDD'):
            raise ValueError('This memory={0} is not is not understandable.. Please use \'RAM\' or \'HDD\'.'.format(memory))
        with _upload(ccMapObj, vmin, vmax) as dm:
            keep = cmh.device_mask(dm, percentile_threshold_no_data, threshold_data_occup)     # :327
            dm.set_mask(keep)
            r = dm.kr(tol, delta, Delta)                                                       # :335-336
            mn, mx = dm.apply_sym(None, masked_mode=0, in_place=False)
            normCCMap = _emit(dm, 1, ccMapObj, ~keep, mn, mx)
        normCCMap.make_unreadable()                                                            # :359-360
        ccMapObj.make_unreadable()
        last_run_info.clear()
        last_run_info.update(method='KR', n=int(keep.size), kept=int(keep.sum()), outer=r['outer'], MVP=r['MVP'],
                             rout=r['rout'], device_ms=r['ms'])
        logger.info(' 	...Finished KR Normalization for {0}: {1} outer iterations, {2} matrix-vector products, '
                    'residual {3:.3e}, {4:.2f} ms on device'.format(_label(ccMapObj), r['outer'], r['MVP'], np.sqrt(r['rout']), r['ms']))
        return _finish(normCCMap, outFile)
    except (KeyboardInterrupt, SystemExit):
        raise
    except Exception as e:
        logger.warning(e)
        logger.warning('Error in normalizing map!!!')
        raise


# ---------------------------------------------------------------------------------------------------
def normalizeCCMapByIC(ccMap, tol=1e-4, vmin=None, vmax=None, outFile=None, iteration=500,
                       percentile_threshold_no_data=None, threshold_data_occup=None, workDir=None):
    """Iterative correction (lib/normalizer.py:505-647).  Returns the normalized CCMAP (matrix left open
    r+, :583), None when ``outFile`` is given, and -- like the reference (:642-647) -- None after logging when
    anything but a device error goes wrong."""
    ccMapObj, ccmapType = cmp.checkCCMapObjectOrFile(ccMap, workDir=workDir)
    logger.info(' Iterative Correction is in process for {0}...'.format(_label(ccMapObj)))
    applyFilter = not (percentile_threshold_no_data is None and threshold_data_occup is None)   # :577-580
    try:
        with _upload(ccMapObj, vmin, vmax) as dm:
            rowsum, zc = dm.row_stats()
            if applyFilter:
                keep = cmh.mask_from_row_stats(rowsum, zc, dm.n, percentile_threshold_no_data, threshold_data_occup)  # :588
                solve_mask = keep
            else:
                # :592-594 -- IC runs on the whole matrix; bNoData = columns that are entirely zero
                keep = zc != dm.n
                solve_mask = np.ones(dm.n, dtype=bool)
            dm.set_mask(solve_mask)
            r = dm.ic(tol, iteration)                                                           # :590 / :593
            mn, mx = dm.apply_sym(None, masked_mode=1, in_place=True)                           # :596-613
            tmap = _emit(dm, 0, ccMapObj, ~keep, mn, mx)
        tmap.make_editable()
        last_run_info.clear()
        last_run_info.update(method='IC', n=int(keep.size), kept=int(keep.sum()), passes=r['passes'], matvecs=r['matvecs'],
                             l1=r['l1'], device_ms=r['ms'], bias=r['bias'])
        logger.info(' 	...Finished Iterative Correction for {0}: {1} passes, last L1 bias change {2:.3e}, '
                    '{3:.2f} ms on device'.format(_label(ccMapObj), r['passes'], r['l1'], r['ms']))
        return _finish(tmap, outFile)
    except (SystemExit, KeyboardInterrupt, GcxError):
        raise
    except Exception as e:
        logger.warning(e)
        logger.warning('Error in Iterative Correction!!!')
        return None


# ---------------------------------------------------------------------------------------------------
def normalizeCCMapByMCFS(ccMap, stats='median', vmin=None, vmax=None, stype='o/e', outFile=None, scaleUpInput=False,
                         percentile_threshold_no_data=None, threshold_data_occup=None, workDir=None):
    """Median/mean contact-frequency scaling = distance-frequency normalization (lib/normalizer.py:742-902).
    Returns the scaled CCMAP (matrix open r+), None with ``outFile``, None after a warning for a bad
    ``stype`` (:810-812) or any non-device error (:897-902)."""
    if stype not in ['o/e', 'o-e']:
        logger.warning('Wrong value {0} for stype: {1}.'.format(stype, ['o/e', 'o-e']))
        return None
    ccMapObj, ccmapType = cmp.checkCCMapObjectOrFile(ccMap, workDir=workDir)
    logger.info(' Median Contact Frequency Scaling is in process for {0}...'.format(_label(ccMapObj)))
    try:
        with _upload(ccMapObj, vmin, vmax) as dm:
            if scaleUpInput:                                                                     # :832-834
                toScale = util.locate_significant_digit_after_decimal(ccMapObj.minvalue) + 1
                dm.scale(10 ** toScale)
            keep = cmh.device_mask(dm, percentile_threshold_no_data, threshold_data_occup)       # :854-855
            avg = dm.diag_stats(stats)                                                           # :858
            mn, mx = dm.diag_apply(None, subtract=(stype == 'o-e'), in_place=False)              # :860-864
            normCCMap = _emit(dm, 1, ccMapObj, ~keep, mn, mx)
        normCCMap.make_editable()
        ccMapObj.make_readable()
        last_run_info.clear()
        last_run_info.update(method='MCFS', n=int(keep.size), stats=stats, stype=stype, avg=avg)
        logger.info(' 	...Finished Median Contact Frequency Scaling for {0}...'.format(_label(ccMapObj)))
        return _finish(normCCMap, outFile)
    except (KeyboardInterrupt, SystemExit, GcxError):
        raise
    except Exception as e:
        logger.warning(e)
        logger.warning('Error in Median Contact Frequency Scaling!!!')
        return None


# ---------------------------------------------------------------------------------------------------
def normalizeCCMapByVCNorm(ccMap, sqroot=False, vmin=None, vmax=None, outFile=None,
                           percentile_threshold_no_data=None, threshold_data_occup=None, workDir=None):
    """Vanilla-coverage normalization (lib/normalizer.py:1041-1168).  Returns the normalized CCMAP (readable),
    None with ``outFile``; non-device errors are printed and give None (:1160-1168)."""
    ccMapObj, ccmapType = cmp.checkCCMapObjectOrFile(ccMap, workDir=workDir)
    try:
        with _upload(ccMapObj, vmin, vmax) as dm:
            keep = cmh.device_mask(dm, percentile_threshold_no_data, threshold_data_occup)       # :1123
            dm.set_mask(keep)
            mn, mx, csum = dm.vc(sqroot=sqroot, in_place=False)                                  # :1128
            normCCMap = _emit(dm, 1, ccMapObj, ~keep, mn, mx)
        normCCMap.make_readable()
        ccMapObj.make_readable()
        last_run_info.clear()
        last_run_info.update(method='VC', n=int(keep.size), kept=int(keep.sum()), sqroot=bool(sqroot))
        return _finish(normCCMap, outFile)
    except (KeyboardInterrupt, SystemExit, GcxError):
        raise
    except Exception as e:
        print('Error in Median Contact Frequency Scaling!!!\n', e)
        return None


# ---------------------------------------------------------------------------------------------------
def _gcmap_unavailable(name):
    def f(*args, **kwargs):
        raise NotImplementedError(
            name + ": the HDF5 .gcmap container (gcMapExplorer/lib/gcmap.py) is outside the accelerated hot path "
            "(SURVEY.md section 8f, row N1) and h5py is not available in this image; convert with the reference's "
            "gcmap tools and call the normalizeCCMapBy* functions per map.")
    f.__name__ = name
    return f


normalizeGCMapByKR = _gcmap_unavailable('normalizeGCMapByKR')
normalizeGCMapByIC = _gcmap_unavailable('normalizeGCMapByIC')
normalizeGCMapByMCFS = _gcmap_unavailable('normalizeGCMapByMCFS')
normalizeGCMapByVCNorm = _gcmap_unavailable('normalizeGCMapByVCNorm')
