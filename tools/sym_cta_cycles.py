"""Per-CTA main-loop cycles of the upper-triangle pass (GCX_SYM_DEBUG=1): where is the load imbalance?"""
import os, sys, ctypes as C
os.environ["GCX_SYM_DEBUG"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from gcmapexplorer_b200 import device as gx, _lib
n = int(sys.argv[1]) if len(sys.argv) > 1 else 24926
with gx.DeviceMap(n) as dm:
    dm.synth(seed=110)
    for rep in range(3):
        dm.bench_matvec(3, 100)
        buf = (C.c_int64 * 1024)(); cnt = C.c_int(0)
        _lib.check(_lib.lib().gcx_debug_sym_cycles(dm._ctx, buf, 1024, C.byref(cnt)))
        cyc = np.array(buf[:cnt.value], dtype=np.float64)
    G = cnt.value
    order = np.argsort(-cyc)
    print("G", G, "min/median/mean/max kcycles: %.0f %.0f %.0f %.0f" % (cyc.min()/1e3, np.median(cyc)/1e3, cyc.mean()/1e3, cyc.max()/1e3))
    print("slowest CTAs:", [(int(c), int(cyc[c]/1e3)) for c in order[:12]])
    print("fastest CTAs:", [(int(c), int(cyc[c]/1e3)) for c in order[-8:]])
    q = np.array_split(np.arange(G), 8)
    print("mean kcycles by CTA-index octile:", [int(cyc[i].mean()/1e3) for i in q])
    print("first 16 CTAs:", [int(x/1e3) for x in cyc[:16]])
