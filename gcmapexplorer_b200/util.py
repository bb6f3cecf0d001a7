"""Host-side helpers of ``gcMapExplorer/lib/util.py`` that the normalization path touches."""
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
