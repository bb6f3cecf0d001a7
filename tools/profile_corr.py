"""Short program for an ncu launch list of the correlation matrix (row N4) at chr1 @ 10 kb."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gcmapexplorer_b200 import device as gx      # noqa: E402
n = int(sys.argv[1]) if len(sys.argv) > 1 else 24926
factor = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0          # e.g. 0.37: a float32-valued non-integer map
with gx.DeviceMap(n) as dm:
    dm.synth(seed=110)
    if factor != 1.0:
        dm.scale(factor)
    rs, zc = dm.row_stats()
    dm.set_mask(rs != 0)
    dm.corr(0.0)
    dm.timer_start()
    dm.corr(0.0)
    print("corr ms", dm.timer_stop(), "route", dm.corr_last_route())
