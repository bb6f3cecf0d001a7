#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29613"
timeout 900 python -m pytest tests/test_gpu_multi.py -x -q > gpurun_out/t_multi9.log 2>&1; echo "multi rc=$?"; tail -15 gpurun_out/t_multi9.log
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "dropin_accepts or fused_apply or mcfs_batch" > gpurun_out/t_quick9.log 2>&1; echo "quick rc=$?"; tail -5 gpurun_out/t_quick9.log
timeout 600 $TR bench.py --gpus 2 --steps 3 --warmup 2 --no-extras > gpurun_out/bench9_n2.json 2> gpurun_out/bench9_n2.err; echo "bench2 rc=$?"; tail -3 gpurun_out/bench9_n2.err
timeout 300 $TR bench.py --workload host_copy_floor --gpus 2 > gpurun_out/floor_n2.json 2> gpurun_out/floor_n2.err; echo "floor rc=$?"; cat gpurun_out/floor_n2.json
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench9_n2.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d["time_to_converge_s"], d["passes"], d["roofline"]["us_per_launch"], d["roofline"]["frac"], d["config"]["parallelism"], d.get("parity"), d["e2e"])
PY
