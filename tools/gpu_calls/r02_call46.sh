#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"k_apply_sym_vc|k_row_stats" -c 2 -o gpurun_out/ncu46_tail -f python bench.py --steps 1 --warmup 0 --no-extras --no-cpu-baseline --no-e2e > gpurun_out/ncu46.log 2>&1; echo "ncu rc=$?"
