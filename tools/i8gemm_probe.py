"""Correctness and speed of the library's own tcgen05 int8 product against NumPy / the data-sheet rate."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from gcmapexplorer_b200 import device as gx
rng = np.random.default_rng(1)
with gx.DeviceMap(8) as dm:
    for mb in (256, 512, 1024):
        L = rng.integers(-127, 128, size=(mb, mb), dtype=np.int8)
        R = rng.integers(-127, 128, size=(mb, mb), dtype=np.int8)
        T, ms = dm.i8gemm(L, R)
        exp = L.astype(np.int32) @ R.astype(np.int32).T
        bad = int((T != exp).sum())
        print("mb %5d: mismatches %d of %d  (%.3f ms)" % (mb, bad, mb * mb, ms), flush=True)
        if bad:
            idx = np.argwhere(T != exp)[:5]
            print("   first bad:", [(int(i), int(j), int(T[i, j]), int(exp[i, j])) for i, j in idx])
            print("   rows with errors:", np.unique(idx[:, 0])[:10], "cols:", np.unique(idx[:, 1])[:10])
        T2, _ = dm.i8gemm(L, L, symmetric=True)
        e2 = L.astype(np.int32) @ L.astype(np.int32).T
        il = np.tril_indices(mb)
        print("   symmetric (lower triangle): mismatches %d" % int((T2[il] != e2[il]).sum()), flush=True)
    for mb in (8192, 24832):
        _, ms = dm.i8gemm(None, None, mb=mb, reps=3, want_result=False)
        print("mb %5d: %.3f ms per product = %.2f POP/s (nominal int8 dense 4.5)" % (mb, ms, 2.0 * mb ** 3 / (ms * 1e-3) / 1e15), flush=True)
        _, ms = dm.i8gemm(None, None, mb=mb, symmetric=True, reps=3, want_result=False)
        print("mb %5d symmetric: %.3f ms" % (mb, ms), flush=True)
