#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "diag or mcfs or median or pooled or statdist or cmstats" > gpurun_out/t28_diag.log 2>&1; echo "diag tests rc=$?"; tail -3 gpurun_out/t28_diag.log
timeout 300 python bench.py --workload wg_25kb_mcfs --steps 5 --warmup 2 --no-cpu-baseline > gpurun_out/bench28_mcfs.json 2> gpurun_out/bench28_mcfs.err; echo "mcfs rc=$?"; cat gpurun_out/bench28_mcfs.json | cut -c1-900
timeout 300 python tools/mcfs_probe.py > gpurun_out/mcfs_probe28.txt 2>&1; echo "probe rc=$?"; tail -25 gpurun_out/mcfs_probe28.txt
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches28_mcfs.csv python bench.py --workload wg_25kb_mcfs --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu28.log 2>&1; echo "ncu rc=$?"
python - <<'PY'
import csv, collections
rows = []
try:
    with open("gpurun_out/launches28_mcfs.csv") as f:
        lines = [l for l in f if not l.startswith("==")]
    for r in csv.DictReader(lines):
        rows.append((r["Kernel Name"], float(r["Metric Value"].replace(",", "")), r.get("Metric Unit", "")))
    # the last step: everything after the last k_mm_init_b
    idx = max(i for i, r in enumerate(rows) if "k_mm_init_b" in r[0])
    for name, v, u in rows[idx:]:
        print("%-60s %12.1f %s" % (name[:60], v, u))
except Exception as e:
    print("launch list unreadable:", e)
PY
