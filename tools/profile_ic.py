"""Short program for ncu: one IC solve at chr1 @ 10 kb (the persistent deferred-scalar pass is ONE launch), the mask pass and the fused tail."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from gcmapexplorer_b200 import device as gx
n = int(sys.argv[1]) if len(sys.argv) > 1 else 24926
with gx.DeviceMap(n) as dm:
    dm.synth(seed=110)
    rs, zc = dm.row_stats()
    dm.set_mask(np.ones(n, dtype=bool))
    r = dm.ic(1e-4, 500)
    print("ic passes", r["passes"], "ms", r["ms"], "us/pass", r["ms"] * 1e3 / r["products_launched"])
    dm.apply_sym_vc(rs != 0, masked_mode=1)
    dm.set_mask(rs != 0)
    k = dm.kr(1e-12, 0.1, 3.0)
    print("kr", k["outer"], k["MVP"], k["ms"])
    print("ok")
