#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 300 python tools/mcfs_probe.py 9970 24926 > gpurun_out/mcfs_probe41.txt 2>&1; echo "probe rc=$?"; grep -v mean gpurun_out/mcfs_probe41.txt | tail -8
GCX_DIAG_APPLY_LANE=0 timeout 300 python tools/mcfs_probe.py 24926 > gpurun_out/mcfs_probe41_f4.txt 2>&1; echo "probe(float4) rc=$?"; grep apply gpurun_out/mcfs_probe41_f4.txt | tail -3
timeout 1500 python -m pytest tests/ -x -q -m gpu > gpurun_out/t41_all.log 2>&1; echo "all gpu tests rc=$?"; tail -2 gpurun_out/t41_all.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench41_n1.json 2> gpurun_out/bench41_n1.err; echo "bench rc=$?"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench41_n1.json").read().strip().splitlines()[-1])
print("value", d["value"], "ms/step", d["ms_per_step"], "frac", d["roofline"]["frac"], "e2e", d["e2e"]["seconds_per_step"])
print("mcfs", d.get("mcfs")); print("kr", json.dumps(d.get("kr"))[:300])
PY
