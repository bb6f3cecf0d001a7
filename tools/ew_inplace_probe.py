"""apply_sym out-of-place vs in-place (same bytes, s = 1 leaves the map unchanged)."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from gcmapexplorer_b200 import device as gx
n = int(sys.argv[1]) if len(sys.argv) > 1 else 24926
with gx.DeviceMap(n) as dm:
    dm.synth(seed=110)
    rs, zc = dm.row_stats()
    dm.set_mask(rs != 0)
    s = np.ones(n)
    out = {}
    for name, inp in (("out_of_place", False), ("in_place", True)):
        dm.apply_sym(s, 0, inp); dm.apply_sym(None, 0, inp)
        dm.profile(True); dm.profile_read(True)
        for _ in range(5):
            dm.apply_sym(None, 0, inp)
        p = dm.profile_read(True)["apply_sym"]
        dm.profile(False)
        ms = p["ms"] / p["launches"]
        out[name] = {"us": round(ms * 1e3, 1), "GBps": round(8.0 * n * n / ms / 1e6, 1)}
    print(json.dumps(out))
