#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29616"
timeout 200 $TR tools/sym_phase_stamps.py 35251 > gpurun_out/stamps26_2gpu.txt 2>&1; grep -v "^\*\|OMP_NUM\|^$\|Warning\|NCCL" gpurun_out/stamps26_2gpu.txt | head -16
timeout 200 $TR tools/sym_phase_stamps.py 49851 > gpurun_out/stamps26_2gpu_b.txt 2>&1; grep -v "^\*\|OMP_NUM\|^$\|Warning\|NCCL" gpurun_out/stamps26_2gpu_b.txt | head -12
