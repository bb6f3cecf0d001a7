#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_diag_hist_b|k_diag_select_b|k_diag_apply_b" -c 5 -o gpurun_out/ncu29_mcfs -f python bench.py --workload wg_25kb_mcfs --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/ncu29.log 2>&1; echo "ncu rc=$?"; tail -3 gpurun_out/ncu29.log
ls -la gpurun_out/ncu29_mcfs.ncu-rep
