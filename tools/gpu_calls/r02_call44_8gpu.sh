#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
TR8="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29644"
TR4="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29645"
timeout 600 $TR8 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/bench44_n8.json 2> gpurun_out/bench44_n8.err; echo "bench8 rc=$?"
timeout 120 $TR8 tools/sym_phase_stamps.py 70501 > gpurun_out/stamps44_8gpu.txt 2>&1; grep -v "^\*\|OMP_NUM\|^$\|Warning\|NCCL" gpurun_out/stamps44_8gpu.txt | head -13
timeout 300 $TR4 bench.py --gpus 4 --steps 5 --warmup 3 --no-extras > gpurun_out/bench44_n4.json 2> gpurun_out/bench44_n4.err; echo "bench4 rc=$?"
python - <<'PY'
import json
for f in ("bench44_n8.json", "bench44_n4.json"):
    try:
        d = json.loads([l for l in open("gpurun_out/" + f) if l.startswith("{")][-1])
        print(f, "value", d["value"], "ms/step", d["ms_per_step"], "ttc", d["time_to_converge_s"], "passes", d["passes"], "us/pass", d["roofline"]["us_per_launch"], "frac", d["roofline"]["frac"], "e2e", d["e2e"]["seconds_per_step"])
        print("   parity", d.get("parity")); print("   chr1_1kb", json.dumps(d.get("chr1_1kb"))[:900])
    except Exception as e:
        print(f, "unreadable", e)
PY
timeout 200 python -m pytest tests/test_gpu_multi.py -x -q > gpurun_out/t44_multi.log 2>&1; echo "multi rc=$?"; tail -1 gpurun_out/t44_multi.log
