#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29643"
timeout 600 $TR bench.py --gpus 2 --steps 5 --warmup 3 --no-extras > gpurun_out/bench43_n2.json 2> gpurun_out/bench43_n2.err; echo "bench2 rc=$?"
timeout 200 $TR tools/sym_phase_stamps.py 35251 > gpurun_out/stamps43_2gpu.txt 2>&1; grep -v "^\*\|OMP_NUM\|^$\|Warning\|NCCL" gpurun_out/stamps43_2gpu.txt | head -12
timeout 300 python -m pytest tests/test_gpu_multi.py -x -q -k "dropin or sharded or symmetry" > gpurun_out/t43_multi.log 2>&1; echo "multi rc=$?"; tail -1 gpurun_out/t43_multi.log
python - <<'PY'
import json
d = json.loads([l for l in open("gpurun_out/bench43_n2.json") if l.startswith("{")][-1])
print("value", d["value"], "ms/step", d["ms_per_step"], "passes", d["passes"], "us/pass", d["roofline"]["us_per_launch"], "frac", d["roofline"]["frac"], "e2e", d["e2e"]["seconds_per_step"])
print("   parity", d.get("parity"))
PY
