#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "deferred or at_size or symmetric_pass or ic_vs_fp64 or ic_golden or kr_golden" > gpurun_out/t_quick23.log 2>&1; echo "quick rc=$?"; tail -3 gpurun_out/t_quick23.log
timeout 120 python tools/ic_quick.py 24926 2>&1 | tee gpurun_out/ic_quick23.txt
for d in 20,32 20,64 30,32 40,64 30,128 40,128; do GCX_SYM_DYN=${d%,*} GCX_SYM_DYN_CHUNK=${d#*,} timeout 120 python tools/ic_quick.py 24926; done 2>&1 | tee -a gpurun_out/ic_quick23.txt
timeout 120 python tools/ic_quick.py 49851 2>&1 | tee -a gpurun_out/ic_quick23.txt
timeout 120 python tools/ic_quick.py 12000 2>&1 | tee -a gpurun_out/ic_quick23.txt
timeout 300 python tools/sym_phase_stamps.py 24926 > gpurun_out/stamps23.txt 2>&1; head -14 gpurun_out/stamps23.txt
