"""Short program for ncu captures of the N3 kernels at chr1 @ 10 kb: transition matrix passes + a few power-iteration steps."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gcmapexplorer_b200 import device as gx      # noqa: E402
n = int(sys.argv[1]) if len(sys.argv) > 1 else 24926
with gx.DeviceMap(n) as dm:
    dm.synth(seed=110)
    rs, zc = dm.row_stats()
    dm.set_mask(rs != 0)
    print("transition min/max", dm.transition(in_place=False))
    dm.swap()
    r = dm.stationary(1e-13, 6)
    print("stationary", r["iterations"], r["delta"], r["ms"])
    dm.invalidate()
    print("median", dm.diag_stats("median")[:3])
