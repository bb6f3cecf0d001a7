#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612"
timeout 120 python tools/ic_quick.py 24926 2>&1 | tee gpurun_out/ic_quick7.txt
timeout 200 $TR tools/ic_quick.py 35251 2>&1 | grep "us per pass" | tee -a gpurun_out/ic_quick7.txt
GCX_PEER_EXCHANGE=0 timeout 200 $TR tools/ic_quick.py 35251 2>&1 | grep "us per pass" | tee -a gpurun_out/ic_quick7.txt
timeout 200 $TR tools/sym_phase_stamps.py 35251 > gpurun_out/stamps7_2gpu.txt 2>&1; grep -v "^\*\|OMP_NUM\|^$" gpurun_out/stamps7_2gpu.txt | head -24
timeout 300 $TR tools/multi_gpu_check.py 9000 balanced > gpurun_out/mg2b.log 2>&1; echo "mg2 rc=$?"
timeout 120 python tools/sym_phase_stamps.py 24926 > gpurun_out/stamps7_n24926.txt 2>&1; head -10 gpurun_out/stamps7_n24926.txt
