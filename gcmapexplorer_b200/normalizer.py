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
    """Stream the CCMAP's float32 matrix into HBM once -- straight from its backing file (parallel pread), from its gzip stream,
    or from the memmap (ccmap.upload_ccmap) -- on one GPU, or row-sharded over several when it does not fit (sharded.py), and
    apply the vmin/vmax clamp there (lib/normalizer.py:299-301, 567-569, 827-829, 1105-1107)."""
    def make_map(n):
        k = sharded.gpus_for_map(n)
        if k > 1:
            # too large for one GPU (or GCX_GPUS=k): row blocks on k GPUs of this node, one host thread each
            return sharded.ShardedDeviceMap(n, devices=list(range(k)))
        return DeviceMap(n)
    dm = cmp.upload_ccmap(ccMapObj, make_map)
    if vmin is not None or vmax is not None:
        dm.clamp(vmin, vmax)
    return dm


def _host_rows(ccMapObj, vmin=None, vmax=None, factor=None):
    """Row i of the map as the device holds it (clamped / scaled), read from the CCMAP's own file -- only ever called for rows with
    negative entries, whose float32 emptiness test (lib/ccmapHelpers.pyx:192) the host repeats in NumPy (ccmapHelpers.float32_empty_rows)."""
    def row(i):
        opened = ccMapObj.matrix is None
        if opened:
            ccMapObj.make_readable()
        try:
            r = np.array(ccMapObj.matrix[i], dtype=np.float32)
        finally:
            if opened:
                ccMapObj.make_unreadable()
        if vmin is not None:
            r[r <= vmin] = 0.0
        if vmax is not None:
            r[r >= vmax] = 0.0
        if factor is not None:
            r = r * np.float32(factor)
        return r
    return row


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
    ccMapObj, ccmapType = cmp.checkCCMapObjectOrFile(ccMap, workDir=workDir, lazy_gz=True)
    logger.info(' KR Normalization will be done through {0}.'.format(memory))
    logger.info(' KR Normalization is in process for {0}...'.format(_label(ccMapObj)))
    delta, Delta = 0.1, 3           # :319-320
    try:
        if memory not in ('RAM', 'HDD'):
            raise ValueError('This memory={0} is not is not understandable.. Please use \'RAM\' or \'HDD\'.'.format(memory))
        with _upload(ccMapObj, vmin, vmax) as dm:
            keep = cmh.device_mask(dm, percentile_threshold_no_data, threshold_data_occup, _host_rows(ccMapObj, vmin, vmax))     # :327
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
    ccMapObj, ccmapType = cmp.checkCCMapObjectOrFile(ccMap, workDir=workDir, lazy_gz=True)
    logger.info(' Iterative Correction is in process for {0}...'.format(_label(ccMapObj)))
    applyFilter = not (percentile_threshold_no_data is None and threshold_data_occup is None)   # :577-580
    try:
        with _upload(ccMapObj, vmin, vmax) as dm:
            rowsum, zc = dm.row_stats()
            if applyFilter:
                keep = cmh.device_mask(dm, percentile_threshold_no_data, threshold_data_occup, _host_rows(ccMapObj, vmin, vmax))  # :588
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
    ccMapObj, ccmapType = cmp.checkCCMapObjectOrFile(ccMap, workDir=workDir, lazy_gz=True)
    logger.info(' Median Contact Frequency Scaling is in process for {0}...'.format(_label(ccMapObj)))
    try:
        with _upload(ccMapObj, vmin, vmax) as dm:
            factor = None
            if scaleUpInput:                                                                     # :832-834
                toScale = util.locate_significant_digit_after_decimal(ccMapObj.minvalue) + 1
                factor = 10 ** toScale
                dm.scale(factor)
            keep = cmh.device_mask(dm, percentile_threshold_no_data, threshold_data_occup, _host_rows(ccMapObj, vmin, vmax, factor))       # :854-855
            avg = dm.diag_stats(stats)                                                           # :858
            mn, mx = dm.diag_apply(None, subtract=(stype == 'o-e'), in_place=False)              # :860-864
            normCCMap = _emit(dm, 1, ccMapObj, ~keep, mn, mx)
        normCCMap.make_editable()
        if ccmapType == 'Object':           # a map loaded from a file inside this call is dropped on return: nothing to re-open
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
    ccMapObj, ccmapType = cmp.checkCCMapObjectOrFile(ccMap, workDir=workDir, lazy_gz=True)
    try:
        with _upload(ccMapObj, vmin, vmax) as dm:
            keep = cmh.device_mask(dm, percentile_threshold_no_data, threshold_data_occup, _host_rows(ccMapObj, vmin, vmax))       # :1123
            dm.set_mask(keep)
            mn, mx, csum = dm.vc(sqroot=sqroot, in_place=False)                                  # :1128
            normCCMap = _emit(dm, 1, ccMapObj, ~keep, mn, mx)
        normCCMap.make_readable()
        if ccmapType == 'Object':
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
# Batch drivers (lib/normalizer.py:397-503, 649-740, 904-1039, 1170-1290): every map of a set, smallest first.
# The set is the HDF5 .gcmap file when h5py is importable, or a map-set directory / list of .ccmap files (mapset.py).
# ---------------------------------------------------------------------------------------------------
def normalizeGCMapByKR(gcMapInputFile, gcMapOutFile, mapSizeCeilingForMemory=20000, vmin=None, vmax=None, tol=1e-12,
                       percentile_threshold_no_data=None, threshold_data_occup=None, compression='lzf', workDir=None, logHandler=None):
    """KR-balance every map of a set (lib/normalizer.py:397-503): maps in ascending size; each result is stored with its
    coarser levels (x2, 'sum', until below 500 bins).  ``mapSizeCeilingForMemory`` picks 'RAM' or 'HDD' in the reference
    (:479-482); both run on the GPU(s) here -- a map too large for one GPU is row-sharded instead."""
    from . import mapset
    ms = mapset.MapSet(gcMapInputFile)
    for mapName in ms.names_smallest_first():
        ccMap = ms.load(mapName, workDir=workDir)
        memory = 'HDD' if ccMap.shape[0] > mapSizeCeilingForMemory else 'RAM'                    # :479-482
        norm_ccmap = normalizeCCMapByKR(ccMap, memory=memory, tol=tol, vmin=vmin, vmax=vmax,
                                        percentile_threshold_no_data=percentile_threshold_no_data,
                                        threshold_data_occup=threshold_data_occup, workDir=workDir)
        if norm_ccmap is not None:
            mapset.add_map(norm_ccmap, gcMapOutFile, compression=compression, generateCoarse=True, coarseningMethod='sum')
        del ccMap
        del norm_ccmap


def normalizeGCMapByIC(gcMapInputFile, gcMapOutFile, vmin=None, vmax=None, tol=1e-12, iteration=500,
                       percentile_threshold_no_data=None, threshold_data_occup=None, compression='lzf', workDir=None, logHandler=None):
    """Iterative correction of every map of a set (lib/normalizer.py:649-740; note the driver's own default tol=1e-12).  A map whose
    normalization fails is skipped with the warnings of normalizeCCMapByIC, like the reference (norm_ccmap is None, :718)."""
    from . import mapset
    ms = mapset.MapSet(gcMapInputFile)
    for mapName in ms.names_smallest_first():
        ccMap = ms.load(mapName, workDir=workDir)
        norm_ccmap = normalizeCCMapByIC(ccMap, tol=tol, iteration=iteration, vmin=vmin, vmax=vmax,
                                        percentile_threshold_no_data=percentile_threshold_no_data,
                                        threshold_data_occup=threshold_data_occup, workDir=workDir)
        if norm_ccmap is not None:
            mapset.add_map(norm_ccmap, gcMapOutFile, compression=compression, generateCoarse=True, coarseningMethod='sum')
        del ccMap
        del norm_ccmap


def _walk_resolutions(ms, gcMapOutFile, compression, workDir, norm_one):
    """MCFS and VC normalize EVERY stored resolution of every map, finest first, and add each without coarsening
    (lib/normalizer.py:1003-1030, 1247-1283); the walk of a map stops at the first resolution that fails."""
    from . import mapset
    for mapName in ms.names_smallest_first():
        for resolution in ms.resolutions(mapName):
            ccMap = ms.load(mapName, resolution=resolution, workDir=workDir)
            norm_ccmap = norm_one(ccMap)
            del ccMap
            if norm_ccmap is None:
                break
            mapset.add_map(norm_ccmap, gcMapOutFile, compression=compression, generateCoarse=False, replaceCMap=False)
            del norm_ccmap


def normalizeGCMapByMCFS(gcMapInputFile, gcMapOutFile, stats='median', vmin=None, vmax=None, stype='o/e', scaleUpInput=False,
                         percentile_threshold_no_data=None, threshold_data_occup=None, compression='lzf', workDir=None, logHandler=None):
    """Distance-frequency scaling of every map of a set at every stored resolution (lib/normalizer.py:904-1039)."""
    from . import mapset
    ms = mapset.MapSet(gcMapInputFile)
    _walk_resolutions(ms, gcMapOutFile, compression, workDir,
                      lambda c: normalizeCCMapByMCFS(c, stats=stats, stype=stype, vmin=vmin, vmax=vmax, scaleUpInput=scaleUpInput,
                                                     percentile_threshold_no_data=percentile_threshold_no_data,
                                                     threshold_data_occup=threshold_data_occup, workDir=workDir))


def normalizeGCMapByVCNorm(gcMapInputFile, gcMapOutFile, sqroot=False, vmin=None, vmax=None, percentile_threshold_no_data=None,
                           threshold_data_occup=None, compression='lzf', workDir=None, logHandler=None):
    """Vanilla-coverage normalization of every map of a set at every stored resolution (lib/normalizer.py:1170-1290)."""
    from . import mapset
    ms = mapset.MapSet(gcMapInputFile)
    _walk_resolutions(ms, gcMapOutFile, compression, workDir,
                      lambda c: normalizeCCMapByVCNorm(c, sqroot=sqroot, vmin=vmin, vmax=vmax,
                                                       percentile_threshold_no_data=percentile_threshold_no_data,
                                                       threshold_data_occup=threshold_data_occup, workDir=workDir))
