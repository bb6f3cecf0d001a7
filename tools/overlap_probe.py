"""Does an async D2H on the copy stream overlap with IC iterations?  Prints wall times of the pieces."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from gcmapexplorer_b200 import device as gx
n = 24926
with gx.DeviceMap(n) as dm:
    dm.synth(seed=110)
    rs, zc = dm.row_stats()
    keep = rs != 0
    ones = np.ones(n, dtype=bool)
    P = len(sys.argv) > 1
    pin = gx.PinnedArray((n, n), pitched=P); pin2 = gx.PinnedArray((n, n), pitched=P)
    print('pitched host buffers:', P)
    dm.set_mask(keep); dm.vc(False, False)
    def t(f):
        dm.sync(); t0 = time.perf_counter(); f(); dm.sync(); dm.download_wait(); return (time.perf_counter() - t0) * 1e3
    dm.set_mask(ones)
    for rep in range(2):
        print("download sync       %.1f ms" % t(lambda: dm.download(1, out=pin.array)))
        print("download async+wait %.1f ms" % t(lambda: (dm.download_async(1, pin2.array), dm.download_wait())))
        print("ic alone            %.1f ms" % t(lambda: dm.ic(1e-4, 500)))
        print("async dl || ic      %.1f ms" % t(lambda: (dm.download_async(1, pin2.array), dm.ic(1e-4, 500), dm.download_wait())))
        print("upload              %.1f ms" % t(lambda: dm.upload(pin.array)))
