"""Mirror of ``gcMapExplorer/lib/ccmapHelpers.pyx`` for the normalization path.

``get_nonzeros_index`` [ccmapHelpers.pyx:156-239]: the O(n^2) part (per-row sum and zero count) runs on the
GPU (kernel k_row_stats); the O(n) percentile / occupancy rules are evaluated on the host with NumPy
itself, which makes the mask bit-identical to the reference by construction.
``remove_zeros`` [:241-301] and ``MemoryMappedArray`` [:43-154] are kept for callers of the inner seam.
"""
import os
import tempfile

import numpy as np

from .device import DeviceMap


class MemoryMappedArray:
    """Temp-file backed ndarray (prefix gcx_nparray_); the file is removed when the object is collected."""

    def __init__(self, shape, fill=None, workDir=None, dtype="float32"):
        self.workDir = workDir or os.environ.get("GCX_WORKDIR") or tempfile.gettempdir()
        self.dtype = dtype
        fd, self.path2matrix = tempfile.mkstemp(prefix="gcx_nparray_", suffix=".tmp", dir=self.workDir, text=False)
        os.close(fd)
        self.arr = np.memmap(self.path2matrix, dtype=dtype, mode="w+", shape=shape)
        if fill is not None:
            self.arr.fill(fill)

    def copy(self):
        other = MemoryMappedArray(self.arr.shape, workDir=self.workDir, dtype=self.dtype)
        other.arr[:] = self.arr[:]
        return other

    def copy_from(self, src):
        self.arr[:] = src.arr[:]

    def copy_to(self, dest):
        dest.arr[:] = self.arr[:]

    def __del__(self):
        try:
            self.arr = None
            if os.path.isfile(self.path2matrix):
                os.remove(self.path2matrix)
        except Exception:
            pass


def mask_from_row_stats(rowsum, raw_zero_count, n, threshold_percentile=None, threshold_data_occup=None,
                        diag_is_zero=None):
    """Host half of get_nonzeros_index: the rules of ccmapHelpers.pyx:190-237 on the device-computed
    per-row sum and zero count.  Returns bData (True = bin has data)."""
    empty = np.asarray(rowsum) == 0.0                                   # :192
    if diag_is_zero is not None:                                        # :196-198 (filterByDiagonal)
        empty = empty | np.asarray(diag_is_zero, dtype=bool)
    zero_count = np.where(empty, 0, np.asarray(raw_zero_count)).astype(np.int64)   # :195, :201
    bData = ~empty
    if threshold_percentile is not None and threshold_data_occup is not None:      # :206-207
        raise AssertionError("Both 'threshold_percentile' and 'threshold_count_ratio' cannot be used simultaneously!")
    if threshold_percentile is not None and zero_count.any():                      # :210-219
        percentile = np.percentile(zero_count[zero_count != 0], threshold_percentile)
        compensation = int((zero_count == 0).sum())
        bData = bData & ~((zero_count - compensation) >= percentile)
    if threshold_data_occup is not None and zero_count.any():                      # :222-227
        bData = bData & ~((zero_count / n) >= (1.0 - threshold_data_occup))
    if bData.sum() == 0:                                                           # :231-237
        if threshold_percentile is not None:
            raise ValueError("All rows contain zero values!!! Try increasing 'threshold_percentile' value or do not use it.")
        if threshold_data_occup is not None:
            raise ValueError("All rows contain zero values!!! Try decreasing 'threshold_data_occup' value or do not use it.")
    return np.asarray(bData, dtype=bool)


def float32_empty_rows(dm, rowsum, host_rows=None):
    """The reference calls a row empty when ``np.sum(row) == 0.0`` in FLOAT32 (lib/ccmapHelpers.pyx:192).  On rows without negative
    entries that is "all entries zero", which the device's float64 row sum decides exactly.  A row WITH negative entries (o-e
    output, correlation maps) can have a float32 sum that cancels to exactly 0 while the float64 sum does not: those rows -- flagged
    by the device -- are re-checked with NumPy's own float32 sum when the host copy is at hand (`host_rows(i)` returns row i as the
    device holds it); without one the float64 rule stands (documented divergence)."""
    empty = np.asarray(rowsum) == 0.0
    if not hasattr(dm, "row_negative_flags"):
        return empty
    suspect = np.nonzero(dm.row_negative_flags() != 0)[0]
    if suspect.size and host_rows is not None:
        for i in suspect:
            empty[i] = bool(np.sum(np.asarray(host_rows(int(i)), dtype=np.float32)) == 0.0)
    return empty


def device_mask(dm, threshold_percentile=None, threshold_data_occup=None, host_rows=None):
    """get_nonzeros_index for a map that is already resident (used by the orchestrators)."""
    rowsum, zc = dm.row_stats()
    if host_rows is not None:
        rowsum = np.where(float32_empty_rows(dm, rowsum, host_rows), 0.0, np.where(np.asarray(rowsum) == 0.0, 1.0, rowsum))
    return mask_from_row_stats(rowsum, zc, dm.n, threshold_percentile, threshold_data_occup)


def get_nonzeros_index(matrix, threshold_percentile=None, threshold_data_occup=None, filterByDiagonal=False):
    """Drop-in for ccmapHelpers.get_nonzeros_index (same signature, same errors, bit-identical result)."""
    with DeviceMap.from_host(matrix) as dm:
        rowsum, zc = dm.row_stats()
        empty = float32_empty_rows(dm, rowsum, lambda i: matrix[i])
        rowsum = np.where(empty, 0.0, np.where(np.asarray(rowsum) == 0.0, 1.0, rowsum))     # carry the float32 verdict
    diag0 = (np.diagonal(matrix) == 0) if filterByDiagonal else None
    return mask_from_row_stats(rowsum, zc, matrix.shape[0], threshold_percentile, threshold_data_occup, diag0)


def remove_zeros(matrix, threshold_percentile=None, threshold_data_occup=None, workDir=None):
    """Drop-in for ccmapHelpers.remove_zeros: (MemoryMappedArray float64 of the kept block, bNoData).  Mask and
    compaction both run on the device from one upload (the reference's element loop, :286-299, is O(n^2) interpreter steps)."""
    with DeviceMap.from_host(matrix) as dm:
        bData = device_mask(dm, threshold_percentile, threshold_data_occup)
        nk = int(bData.sum())
        A = MemoryMappedArray((nk, nk), workDir=workDir, dtype="float64")
        if isinstance(A.arr, np.ndarray) and A.arr.flags.c_contiguous:
            dm.compact_f64(bData, out=A.arr)
        else:
            A.arr[:] = dm.compact_f64(bData)
    return A, ~bData
