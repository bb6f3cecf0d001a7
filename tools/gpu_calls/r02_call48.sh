#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches48_bench.csv python bench.py --steps 2 --warmup 1 --no-extras --no-cpu-baseline --no-e2e > gpurun_out/ncu48.log 2>&1; echo "ncu rc=$?"
tail -c 400 gpurun_out/ncu48.log
