#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29615"
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/t_gpu24.log 2>&1; echo "gpu rc=$?"; tail -4 gpurun_out/t_gpu24.log
timeout 200 $TR tools/ic_quick.py 35251 2>&1 | grep "us per pass" | tee gpurun_out/ic_quick24.txt
timeout 120 python tools/ic_quick.py 24926 2>&1 | tee -a gpurun_out/ic_quick24.txt
timeout 120 python tools/ic_quick.py 12000 2>&1 | tee -a gpurun_out/ic_quick24.txt
timeout 300 $TR tools/multi_gpu_check.py 30000 balanced > gpurun_out/mg24.log 2>&1; echo "mg rc=$?"; tail -1 gpurun_out/mg24.log | cut -c1-300
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench24_n1.json 2> gpurun_out/bench24_n1.err; echo "bench rc=$?"
timeout 600 $TR bench.py --gpus 2 --steps 5 --warmup 3 --no-extras > gpurun_out/bench24_n2.json 2> gpurun_out/bench24_n2.err; echo "bench2 rc=$?"
python - <<'PY'
import json
for f in ("bench24_n1.json", "bench24_n2.json"):
    d = json.loads(open("gpurun_out/" + f).read().strip().splitlines()[-1])
    print(f, "value", d["value"], "ms/step", d["ms_per_step"], "ttc", d["time_to_converge_s"], "passes", d["passes"], "us/pass", d["roofline"]["us_per_launch"], "frac", d["roofline"]["frac"], "e2e", d["e2e"]["seconds_per_step"], d.get("parity", {}).get("ok"))
    if "kr" in d: print("  kr", d["kr"]["time_to_converge_s"], d["kr"]["frac"], " mcfs", d["mcfs"]["ms_per_step"], " other", d["roofline"]["other_kernels_ms_per_step"])
PY
