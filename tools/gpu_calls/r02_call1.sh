#!/bin/bash
# round 2, GPU call 1: validate the deferred-scalar IC pass, generate the at-size oracle fixtures, full GPU suite, default bench
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv > gpurun_out/gpu.txt; nproc >> gpurun_out/gpu.txt; free -g >> gpurun_out/gpu.txt
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "deferred or ic_vs_fp64 or symmetric_pass or ic_golden" > gpurun_out/t_quick.log 2>&1; echo "quick rc=$?" >> gpurun_out/t_quick.log
tail -5 gpurun_out/t_quick.log
timeout 900 python tests/golden/gen_golden_large.py > gpurun_out/gen_large.log 2>&1; echo "gen rc=$?" >> gpurun_out/gen_large.log
mkdir -p gpurun_out/large && cp tests/golden/large/*.npz gpurun_out/large/ 2>/dev/null
tail -8 gpurun_out/gen_large.log
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/t_gpu.log 2>&1; echo "gpu rc=$?" >> gpurun_out/t_gpu.log
tail -5 gpurun_out/t_gpu.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc=$?"
GCX_IC_DEFER=0 timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --no-extras > gpurun_out/bench_n1_literal.json 2> gpurun_out/bench_n1_literal.err; echo "bench literal rc=$?"
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"
tail -c 1500 gpurun_out/bench_n1.json
