#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/t_gpu15.log 2>&1; echo "gpu rc=$?"; tail -6 gpurun_out/t_gpu15.log
timeout 300 python tools/io_probe.py > gpurun_out/io_probe15.txt 2>&1; echo "io rc=$?"; tail -1 gpurun_out/io_probe15.txt
timeout 300 python tools/dropin_e2e_probe.py > gpurun_out/dropin15.txt 2>&1; tail -1 gpurun_out/dropin15.txt
python - <<'PY'
# per-kernel time of the batched MCFS step (CUDA events per class) for the 24 whole-genome maps
import os, sys, time
sys.path.insert(0, os.getcwd())
import numpy as np
import bench
from gcmapexplorer_b200 import device as gx
maps = []
for k, n in enumerate(bench.WG_25KB_BINS):
    d = gx.DeviceMap(n); d.synth(seed=25000 + k); maps.append(d)
b = gx.MapBatch(maps)
for stats in ("median", "mean"):
    for _ in range(3): b.mcfs(stats)
    maps[0].profile(True); maps[0].profile_read(True)
    t0 = time.perf_counter()
    for _ in range(5): b.mcfs(stats)
    for d in maps: d.sync()
    wall = (time.perf_counter() - t0) / 5
    p = maps[0].profile_read(True); maps[0].profile(False)
    print(stats, "wall ms/step %.3f" % (wall * 1e3), {k: (v["launches"] // 5, round(v["ms"] / 5, 3)) for k, v in p.items() if v["launches"]})
PY
