#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
for v in 0 1 2 3 4; do echo "variant $v"; GCX_I8_VARIANT=$v timeout 200 python tools/i8gemm_probe.py 2>&1 | grep "mb  8192\|mb 24832\|mismatches" | tail -5; done 2>&1 | tee gpurun_out/i8_variants.txt
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "corr or own_int8 or correlation" > gpurun_out/t_corr18.log 2>&1; echo "corr tests rc=$?"; tail -4 gpurun_out/t_corr18.log
timeout 200 python tools/profile_corr.py 2>&1 | tail -1
GCX_CORR_SELFVAR=0 timeout 200 python tools/profile_corr.py 2>&1 | tail -1
