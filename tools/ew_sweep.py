"""Times the elementwise kernels (apply_sym, vc_apply, diag_apply) for the current GCX_EW_VARIANT."""
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
    for _ in range(2):
        dm.apply_sym(s, 0, False); dm.vc(False, False); dm.diag_apply(s, False, False)
    dm.profile(True); dm.profile_read(True)
    for _ in range(5):
        dm.apply_sym(None, 0, False); dm.vc(False, False); dm.diag_apply(None, False, False)
    p = dm.profile_read(True)
    out = {"variant": os.environ.get("GCX_EW_VARIANT", "0"), "n": n}
    for k in ("apply_sym", "vc_apply", "diag_apply"):
        ms = p[k]["ms"] / p[k]["launches"]
        out[k] = {"us": round(ms * 1e3, 1), "GBps": round(8.0 * n * n / ms / 1e6, 1)}
    print(json.dumps(out))
