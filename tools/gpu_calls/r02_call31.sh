#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "diag or mcfs or median or pooled or statdist or cmstats" > gpurun_out/t31_diag.log 2>&1; echo "diag tests rc=$?"; tail -3 gpurun_out/t30_diag.log
timeout 300 python bench.py --workload wg_25kb_mcfs --steps 5 --warmup 2 --no-cpu-baseline > gpurun_out/bench31_mcfs.json 2> gpurun_out/bench31_mcfs.err; echo "mcfs rc=$?"; cut -c1-200 gpurun_out/bench31_mcfs.json
timeout 300 python tools/mcfs_probe.py 9970 24926 > gpurun_out/mcfs_probe31.txt 2>&1; echo "probe rc=$?"; tail -5 gpurun_out/mcfs_probe31.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_diag_hist_b|k_diag_select_b|k_diag_apply_b" -c 9 -o gpurun_out/ncu31_mcfs -f python bench.py --workload wg_25kb_mcfs --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/ncu31.log 2>&1; echo "ncu rc=$?"
