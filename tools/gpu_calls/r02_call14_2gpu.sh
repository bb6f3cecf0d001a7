#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29614"
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "deferred or at_size or symmetric_pass or ic_vs_fp64 or ic_golden or kr_golden or hic_norm or streaming or batch_driver or mcfs_bad" > gpurun_out/t_quick14.log 2>&1; echo "quick rc=$?"; tail -4 gpurun_out/t_quick14.log
timeout 900 python -m pytest tests/test_gpu_multi.py -x -q > gpurun_out/t_multi14.log 2>&1; echo "multi rc=$?"; tail -3 gpurun_out/t_multi14.log
timeout 120 python tools/ic_quick.py 24926 2>&1 | tee gpurun_out/ic_quick14.txt
timeout 200 $TR tools/ic_quick.py 35251 2>&1 | grep "us per pass" | tee -a gpurun_out/ic_quick14.txt
timeout 200 $TR tools/ic_quick.py 49851 2>&1 | grep "us per pass" | tee -a gpurun_out/ic_quick14.txt
timeout 120 python tools/ic_quick.py 49851 2>&1 | tee -a gpurun_out/ic_quick14.txt
timeout 300 $TR tools/multi_gpu_check.py 49000 balanced > gpurun_out/mg14.log 2>&1; echo "mg rc=$?"; tail -1 gpurun_out/mg14.log | cut -c1-400
timeout 300 python tools/io_probe.py > gpurun_out/io_probe14.txt 2>&1; echo "io rc=$?"; tail -1 gpurun_out/io_probe14.txt
