"""Mirror of ``gcMapExplorer/lib/cmstats.py::getAvgContactByDistance`` (cmstats.py:522-607).
The rest of the reference module (map-vs-map correlation) is out of scope."""
import numpy as np

from . import ccmap as cmp
from .device import DeviceMap


def getAvgContactByDistance(ccmaps, stats="median", removeOutliers=False, outliersThreshold=3.5):
    """Per-diagonal median (default) or mean of the non-zero entries, float64[n]; 0.0 for a diagonal with
    no non-zero entry.  One CCMAP (what the normalizer passes, lib/normalizer.py:858) is supported on the
    GPU; pooled lists and removeOutliers=True are not on the normalization path and raise."""
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
    if len(maps) != 1:
        raise NotImplementedError("pooling several maps is not on the normalization hot path (SURVEY.md 8a, a7)")
    m = maps[0]
    if m.matrix is None:
        raise ValueError("CCMAP is not readable: call make_readable() first")
    with DeviceMap.from_host(m.matrix) as dm:
        return dm.diag_stats(stats)
