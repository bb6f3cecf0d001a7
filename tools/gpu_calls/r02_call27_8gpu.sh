#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
TR8="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29631"
TR4="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29632"
timeout 900 $TR8 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/bench27_n8.json 2> gpurun_out/bench27_n8.err; echo "bench8 rc=$?"
timeout 600 $TR4 bench.py --gpus 4 --steps 5 --warmup 3 --no-extras > gpurun_out/bench27_n4.json 2> gpurun_out/bench27_n4.err; echo "bench4 rc=$?"
timeout 300 $TR8 tools/sym_phase_stamps.py 70501 > gpurun_out/stamps27_8gpu.txt 2>&1; grep -v "^\*\|OMP_NUM\|^$\|Warning\|NCCL" gpurun_out/stamps27_8gpu.txt | head -14
timeout 600 python -m pytest tests/test_gpu_multi.py -x -q > gpurun_out/t_multi27.log 2>&1; echo "multi rc=$?"; tail -2 gpurun_out/t_multi27.log
python - <<'PY'
import json
for f in ("bench27_n8.json", "bench27_n4.json"):
    try:
        d = json.loads(open("gpurun_out/" + f).read().strip().splitlines()[-1])
        print(f, "value", d["value"], "ms/step", d["ms_per_step"], "ttc", d["time_to_converge_s"], "passes", d["passes"], "us/pass", d["roofline"]["us_per_launch"], "frac", d["roofline"]["frac"], "e2e", d["e2e"]["seconds_per_step"])
        print("   parity", d.get("parity")); print("   chr1_1kb", d.get("chr1_1kb"))
    except Exception as e:
        print(f, "unreadable", e)
PY
