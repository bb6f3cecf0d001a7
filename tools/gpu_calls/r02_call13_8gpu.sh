#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo8.txt 2>&1
TR8="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29621"
TR4="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29622"
timeout 900 $TR8 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/bench13_n8.json 2> gpurun_out/bench13_n8.err; echo "bench8 rc=$?"; tail -2 gpurun_out/bench13_n8.err
timeout 300 $TR8 bench.py --workload host_copy_floor --gpus 8 > gpurun_out/floor_n8.json 2> gpurun_out/floor_n8.err; echo "floor8 rc=$?"; cat gpurun_out/floor_n8.json
timeout 300 $TR8 bench.py --workload host_copy_floor --gpus 8 --no-numa-bind > gpurun_out/floor_n8_unbound.json 2> /dev/null; cat gpurun_out/floor_n8_unbound.json
timeout 300 $TR8 tools/ic_quick.py 70501 2>&1 | grep "us per pass" | tee gpurun_out/ic_quick13.txt
GCX_PEER_EXCHANGE=0 timeout 300 $TR8 tools/ic_quick.py 70501 2>&1 | grep "us per pass" | tee -a gpurun_out/ic_quick13.txt
timeout 300 $TR8 tools/sym_phase_stamps.py 70501 > gpurun_out/stamps13_8gpu.txt 2>&1; grep -v "^\*\|OMP_NUM\|^$\|Warning" gpurun_out/stamps13_8gpu.txt | head -16
timeout 600 $TR4 bench.py --gpus 4 --steps 5 --warmup 3 --no-extras > gpurun_out/bench13_n4.json 2> gpurun_out/bench13_n4.err; echo "bench4 rc=$?"
timeout 300 $TR8 tools/multi_gpu_check.py 30000 balanced > gpurun_out/mg8.log 2>&1; echo "mg8 rc=$?"; tail -1 gpurun_out/mg8.log | cut -c1-600
python - <<'PY'
import json
for f in ("bench13_n8.json", "bench13_n4.json"):
    try:
        d = json.loads(open("gpurun_out/" + f).read().strip().splitlines()[-1])
        print(f, "value", d["value"], "ms/step", d["ms_per_step"], "ttc", d["time_to_converge_s"], "passes", d["passes"], "us/pass", d["roofline"]["us_per_launch"], "frac", d["roofline"]["frac"])
        print("   e2e", d["e2e"]["seconds_per_step"], d["e2e"]["host_copy_gbs_per_gpu"], d["e2e"]["numa"]); print("   parity", d.get("parity")); print("   chr1_1kb", d.get("chr1_1kb"))
    except Exception as e:
        print(f, "unreadable", e)
PY
