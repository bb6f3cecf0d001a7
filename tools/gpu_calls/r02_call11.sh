#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/t_gpu11.log 2>&1; echo "gpu rc=$?"; tail -12 gpurun_out/t_gpu11.log
timeout 300 python tools/dropin_e2e_probe.py > gpurun_out/dropin11.txt 2>&1; echo "dropin rc=$?"; tail -3 gpurun_out/dropin11.txt
