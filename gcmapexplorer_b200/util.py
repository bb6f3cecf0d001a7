"""Host-side helpers of ``gcMapExplorer/lib/util.py`` that the normalization path touches."""
import re

import numpy as np


def locate_significant_digit_after_decimal(value):
    """Number of leading zeros after the decimal point of a small number (lib/util.py:53-73);
    used by scaleUpInput (lib/normalizer.py:832-834)."""
    location = 0
    while int(abs(value) * 10 ** location) <= 0:
        location += 1
    return location


def kth_diag_indices(k, a):
    """Index arrays of the k-th diagonal of square array `a` (lib/util.py:75-98).  On the device this is
    plain |i-j| arithmetic; kept for API compatibility."""
    rows, cols = np.diag_indices_from(a)
    if k < 0:
        return rows[:k], cols[-k:]
    if k > 0:
        return rows[k:], cols[:-k]
    return rows, cols


def detectOutliers1D(points, thresh=3.5):
    """Boolean array, True where the modified z-score (0.6745 * |x - median| / MAD, Iglewicz & Hoaglin) exceeds
    `thresh` (lib/util.py:100-138).  O(n) work on a vector: stays on the host."""
    pts = np.asarray(points)
    if pts.ndim == 1:
        pts = pts[:, None]
    dist = np.sqrt(np.sum((pts - np.median(pts, axis=0)) ** 2, axis=-1))
    return 0.6745 * dist / np.median(dist) > thresh


def detectOutliersMasked1D(points, thresh=3.5):
    """Masked-array variant (lib/util.py:140-174): outliers among the unmasked points are masked too; the result
    is a masked array of the input's shape in which everything masked reads 0."""
    values = points.compressed().copy()
    values[detectOutliers1D(values, thresh=thresh)] = 0.0
    full = np.zeros(points.shape, dtype=points.dtype)
    full[~np.ma.getmaskarray(points)] = values
    return np.ma.masked_values(full, 0.0)


_UNIT_FACTORS = (('gb', 10 ** 9), ('mb', 10 ** 6), ('kb', 10 ** 3))


def resolutionToBinsize(resolution):
    """'10kb' -> 10000, '1.6mb' -> 1600000, '160b' -> 160 (lib/util.py:176-231); ValueError for an unknown unit."""
    factor = None
    if re.search(r'\d+b', resolution) is not None:
        factor = 1
    else:
        for unit, f in reversed(_UNIT_FACTORS):
            if unit in resolution:
                factor = f
                break
    if factor is None:
        raise ValueError(" '{0}' keyword of resolution not implemented.".format(resolution))
    m = re.search(r'(\d+.\d+)', resolution)          # the reference's pattern: '.' matches any character
    base = float(m.group()) if m is not None else int(re.search(r'(\d+)', resolution).group())
    return int(base * factor)


def binsizeToResolution(binsize):
    """10000 -> '10kb', 125500 -> '125.5kb', 1634300 -> '1.6343mb' (lib/util.py:233-286)."""
    units = ('b', 'kb', 'mb', 'gb')
    power = 0
    value = binsize
    while value >= 1000:
        power += 1
        value = binsize / (1000 ** power)
    text = str(int(value)) if value == int(value) else str(value)
    return text + units[power]
