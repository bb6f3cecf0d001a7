#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/t_gpu12.log 2>&1; echo "gpu rc=$?"; tail -6 gpurun_out/t_gpu12.log
timeout 300 python tools/io_probe.py > gpurun_out/io_probe12.txt 2>&1; echo "io rc=$?"; tail -2 gpurun_out/io_probe12.txt
