#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo2.txt 2>&1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611"
timeout 300 $TR tools/multi_gpu_check.py 9000 balanced > gpurun_out/mg2_balanced.log 2>&1; echo "mg2 balanced rc=$?"; tail -3 gpurun_out/mg2_balanced.log
timeout 300 $TR tools/multi_gpu_check.py 12345 balanced > gpurun_out/mg2_balanced_b.log 2>&1; echo "mg2 balanced b rc=$?"; tail -2 gpurun_out/mg2_balanced_b.log
GCX_PEER_EXCHANGE=0 timeout 300 $TR tools/multi_gpu_check.py 9000 balanced > gpurun_out/mg2_balanced_nccl.log 2>&1; echo "mg2 nccl rc=$?"; tail -2 gpurun_out/mg2_balanced_nccl.log
timeout 300 $TR tools/multi_gpu_check.py 6001 rows > gpurun_out/mg2_rows.log 2>&1; echo "mg2 rows rc=$?"; tail -2 gpurun_out/mg2_rows.log
timeout 600 python -m pytest tests/test_gpu_multi.py -x -q > gpurun_out/t_multi2.log 2>&1; echo "multi rc=$?"; tail -3 gpurun_out/t_multi2.log
timeout 600 $TR bench.py --gpus 2 --steps 3 --warmup 2 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo "bench2 rc=$?"; tail -c 1800 gpurun_out/bench_n2.json; tail -5 gpurun_out/bench_n2.err
GCX_PEER_EXCHANGE=0 timeout 600 $TR bench.py --gpus 2 --steps 3 --warmup 2 --no-e2e --no-extras --no-parity > gpurun_out/bench_n2_nccl.json 2> gpurun_out/bench_n2_nccl.err; echo "bench2 nccl rc=$?"
python - <<'PY'
import json
for f in ("bench_n2.json", "bench_n2_nccl.json"):
    try:
        d = json.loads(open("gpurun_out/" + f).read().strip().splitlines()[-1])
        print(f, d["value"], d["ms_per_step"], d["time_to_converge_s"], d["passes"], d["roofline"]["us_per_launch"], d["roofline"]["frac"], d["config"]["parallelism"], d.get("parity"))
    except Exception as e:
        print(f, "unreadable", e)
PY
