"""Short program for ncu captures: a few launches of each matrix-pass kernel at chr1 @ 10 kb."""
import sys
sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))
import numpy as np
from gcmapexplorer_b200 import device as gx
n = int(sys.argv[1]) if len(sys.argv) > 1 else 24926
variant = int(sys.argv[2]) if len(sys.argv) > 2 else -1
with gx.DeviceMap(n) as dm:
    dm.synth(seed=110)
    rs, zc = dm.row_stats()
    print("matvec us", dm.bench_matvec(4, variant) * 1e3)
    dm.set_mask(rs != 0)
    dm.apply_sym(np.ones(n), masked_mode=0, in_place=False)
    dm.vc(False, in_place=False)
    dm.diag_apply(np.ones(n), subtract=False, in_place=False)
    dm.set_mask(np.ones(n, dtype=bool))
    r = dm.ic(1e-4, 6)
    print("ic sym", r["symmetric"], r["passes"])
    print("sym matvec us", dm.bench_matvec(4, 100) * 1e3)
    print("ok")
