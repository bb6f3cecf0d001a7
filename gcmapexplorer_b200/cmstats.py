"""Mirror of ``gcMapExplorer/lib/cmstats.py::getAvgContactByDistance`` (cmstats.py:522-607), single maps and pooled lists.
The rest of the reference module (map-vs-map correlation) is out of scope."""
import numpy as np

from . import ccmap as cmp
from .device import DeviceMap


def getAvgContactByDistance(ccmaps, stats="median", removeOutliers=False, outliersThreshold=3.5):
    """Per-diagonal median (default) or mean of the non-zero entries, float64[n]; 0.0 for a diagonal with no non-zero entry.
    One CCMAP (what the normalizer passes, lib/normalizer.py:858) or a list of them: the d-th diagonals of all maps are pooled
    (lib/cmstats.py:572-578; n = bins of the largest map) -- one set of per-diagonal histograms fed by every map on the GPU.
    removeOutliers=True (a modified z-score filter per diagonal, never requested by the normalizers) is not implemented."""
    if cmp.is_ccmap(ccmaps):
        maps = [ccmaps]
    elif isinstance(ccmaps, list):
        for m in ccmaps:
            if not cmp.is_ccmap(m):
                raise TypeError("Not a list of CCMAP instances")
        maps = ccmaps
    else:
        raise TypeError("Not a CCMAP instance or a list of CCMAP instances")
    if removeOutliers:
        raise NotImplementedError("removeOutliers=True is not on the normalization hot path (SURVEY.md 8a, a7)")
    for m in maps:
        if m.matrix is None:
            raise ValueError("CCMAP is not readable: call make_readable() first")
    if len(maps) == 1:
        with DeviceMap.from_host(maps[0].matrix) as dm:
            return dm.diag_stats(stats)
    from .device import MapBatch
    dms = []
    try:
        for m in maps:
            dms.append(DeviceMap.from_host(m.matrix))
        return MapBatch(dms).diag_stats_pooled(stats)
    finally:
        for d in dms:
            d.close()
