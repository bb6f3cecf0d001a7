"""Time the per-diagonal statistics (MCFS) at a few sizes; run under ncu for per-launch times."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gcmapexplorer_b200 import device as gx      # noqa: E402


def main():
    sizes = [int(a) for a in sys.argv[1:]] or [2500, 9970, 24926]
    for n in sizes:
        dm = gx.DeviceMap(n)
        dm.synth(seed=3, scale=200.0, alpha=1.0, empty_frac=0.01)
        for stats in ("mean", "median"):
            dm.diag_stats(stats)
            reps = 5
            dm.timer_start()
            t0 = time.perf_counter()
            for _ in range(reps):
                avg = dm.diag_stats(stats)
            ms = dm.timer_stop() / reps
            wall = (time.perf_counter() - t0) / reps * 1e3
            print(f"n={n} {stats}: {ms:.3f} ms device, {wall:.3f} ms wall, lower-triangle bytes {2.0 * n * n / 1e9:.3f} GB"
                  f" -> {2.0 * n * n / 1e9 / (ms * 1e-3):.0f} GB/s per traversal-equivalent, checksum {float(np.sum(avg)):.6e}")
        for sub in (False, True):
            dm.diag_apply(None, subtract=sub, in_place=False)
            dm.timer_start()
            for _ in range(5):
                dm.diag_apply(None, subtract=sub, in_place=False)
            ms = dm.timer_stop() / 5
            print(f"n={n} apply ({'o-e' if sub else 'o/e'}): {ms:.3f} ms device -> {8.0 * n * n / 1e9 / (ms * 1e-3):.0f} GB/s on 8 n^2 bytes")
        del dm


if __name__ == "__main__":
    main()
