#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 300 python tools/sym_phase_stamps.py 24926 > gpurun_out/stamps_n24926.txt 2>&1; echo "stamps rc=$?"
cat gpurun_out/stamps_n24926.txt
timeout 600 python tools/sym_sweep.py 24926 > gpurun_out/sym_sweep.txt 2>&1; echo "sweep rc=$?"
cat gpurun_out/sym_sweep.txt
for v in 0 1 2 3 4; do GCX_EW2_VARIANT=$v timeout 120 python - <<'PY' >> gpurun_out/ew2.txt 2>&1
import os, numpy as np, sys
sys.path.insert(0, os.getcwd())
from gcmapexplorer_b200 import device as gx
n = 24926
with gx.DeviceMap(n) as dm:
    dm.synth(seed=110)
    rs, _ = dm.row_stats()
    ones = np.ones(n, dtype=bool)
    dm.set_mask(ones)
    dm.ic(1e-4, 500)
    dm.profile(True); dm.profile_read(True)
    for _ in range(6):
        a, b = dm.apply_sym_vc(rs != 0, masked_mode=1)
    p = dm.profile_read(True)
    us = p["apply_sym"]["ms"] / p["apply_sym"]["launches"] * 1e3
    c2, c3 = dm.checksum(1), dm.checksum(2)
    dm.apply_sym(None, masked_mode=1, in_place=False); d2 = dm.checksum(1)
    dm.set_mask(rs != 0); dm.vc(False, in_place=False); d3 = dm.checksum(1)
    print("EW2 variant", os.environ.get("GCX_EW2_VARIANT"), "fused apply+VC: %.1f us = %.0f GB/s on 12 n^2; identical to the two passes: %s %s" % (us, 12.0 * n * n / us / 1e3, c2 == d2, c3 == d3), a, b)
PY
done
cat gpurun_out/ew2.txt
