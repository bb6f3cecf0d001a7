"""Sweep of the scheduling knobs of the upper-triangle pass (share of the units in the dynamic pool, chunk length, static skew)
at chr1 @ 10 kb: microseconds per fused IC launch (CUDA events over a whole run / launches that ran a product)."""
import os, sys, itertools
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from gcmapexplorer_b200 import device as gx
n = int(sys.argv[1]) if len(sys.argv) > 1 else 24926
grid = [(d, c) for d in (0, 5, 10, 20, 35, 50) for c in (4, 8, 16, 32)] if len(sys.argv) < 3 else [tuple(map(int, a.split(","))) for a in sys.argv[2:]]
ones = np.ones(n, dtype=bool)
for dyn, chunk in grid:
    if dyn == 0 and chunk != 4:
        continue
    os.environ["GCX_SYM_DYN"] = str(dyn)
    os.environ["GCX_SYM_DYN_CHUNK"] = str(chunk)
    with gx.DeviceMap(n) as dm:
        dm.synth(seed=110)
        dm.set_mask(ones)
        best = 1e9
        for rep in range(4):
            r = dm.ic(1e-4, 500)
            best = min(best, r["ms"] * 1e3 / r["products_launched"])
        print("dyn %3d%% chunk %2d bands: %.2f us per launch (passes %d, bytes/launch %.4f GB)" % (dyn, chunk, best, r["passes"], r["pass_bytes"] / 1e9), flush=True)
