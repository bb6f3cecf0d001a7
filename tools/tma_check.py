"""Correctness of a matvec variant vs variant 2 (LDG path) through a full IC run at small sizes."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from gcmapexplorer_b200 import device as gx
v = sys.argv[1]
ok = True
for n in (5, 97, 1031, 3000):
    res = []
    for var in ("2", v):
        os.environ["GCX_MATVEC_VARIANT"] = var
        os.environ["GCX_IC_FUSED"] = "0"
        with gx.DeviceMap(n) as dm:
            dm.synth(seed=3)
            dm.set_mask(np.ones(n, dtype=bool))
            res.append(dm.ic(1e-4, 500))
    same = res[0]["passes"] == res[1]["passes"] and np.allclose(res[0]["bias"], res[1]["bias"], rtol=1e-12)
    print(n, v, res[0]["passes"], res[1]["passes"], "OK" if same else "MISMATCH")
    ok &= same
sys.exit(0 if ok else 1)
