"""Where does a launch of the deferred-scalar IC pass spend its time?  Per-CTA globaltimer stamps of one launch (GCX_SYM_DEBUG=1):
entry -> first tiles issued -> vector phase -> grid barrier -> product (static share, dynamic chunks) -> end."""
import os, sys, ctypes as C
os.environ["GCX_SYM_DEBUG"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from gcmapexplorer_b200 import device as gx, _lib
n = int(sys.argv[1]) if len(sys.argv) > 1 else 24926
with gx.DeviceMap(n) as dm:
    dm.synth(seed=110)
    dm.set_mask(np.ones(n, dtype=bool))
    for rep in range(2):
        r = dm.ic(1e-4, 500)
    buf = (C.c_int64 * (16 * 1024))(); cnt = C.c_int(0)
    _lib.check(_lib.lib().gcx_debug_sym_stamps(dm._ctx, buf, 16 * 1024, C.byref(cnt)))
    G = cnt.value
    st = np.array(buf[:16 * G], dtype=np.int64).reshape(G, 16)
    t0 = st[:, 0].min()
    names = ["entry", "tiles issued", "vector phase done", "past grid barrier", "product starts", "static share done", "all done"]
    print("n", n, "passes", r["passes"], "device ms", r["ms"], "-> us per launch", r["ms"] * 1e3 / r["products_launched"], "G", G)
    for k, nm in enumerate(names):
        v = (st[:, k] - t0) / 1e3
        print("%-20s min %8.2f  p10 %8.2f  median %8.2f  p90 %8.2f  max %8.2f us" % (nm, v.min(), np.percentile(v, 10), np.median(v), np.percentile(v, 90), v.max()))
    ch = st[:, 7]
    print("dynamic chunks per CTA: min %d median %d max %d total %d" % (ch.min(), np.median(ch), ch.max(), ch.sum()))
    end = (st[:, 6] - t0) / 1e3
    stat = (st[:, 5] - t0) / 1e3
    print("static done by SM half (CTA < 148 / >= 148): %.1f / %.1f ; end: %.1f / %.1f" % (stat[:148].mean(), stat[148:].mean(), end[:148].mean(), end[148:].mean()))
    order = np.argsort(-end)
    print("last CTAs to finish:", [(int(c), round(float(end[c]), 1), int(ch[c])) for c in order[:10]])
    print("first CTAs to finish:", [(int(c), round(float(end[c]), 1), int(ch[c])) for c in order[-6:]])
