#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests/ -x -q -m gpu > gpurun_out/t45_all.log 2>&1; echo "all gpu tests rc=$?"; tail -1 gpurun_out/t45_all.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench45_n1.json 2> gpurun_out/bench45_n1.err; echo "bench rc=$?"
python - <<'PY'
import json
d = json.loads([l for l in open("gpurun_out/bench45_n1.json") if l.startswith("{")][-1])
print("value", d["value"], "ms/step", d["ms_per_step"], "frac", d["roofline"]["frac"], "e2e", d["e2e"]["seconds_per_step"], "mcfs", d["mcfs"]["ms_per_step"], "kr", d["kr"]["time_to_converge_s"], "launches", d["gpu_launches"])
PY
