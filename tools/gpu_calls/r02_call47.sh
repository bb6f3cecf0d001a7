#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "fused or apply or vc or VC or dropin or golden" > gpurun_out/t47.log 2>&1; echo "tests rc=$?"; tail -1 gpurun_out/t47.log
timeout 300 python bench.py --steps 5 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/bench47_n1.json 2> gpurun_out/bench47_n1.err; echo "bench rc=$?"
python - <<'PY'
import json
d = json.loads([l for l in open("gpurun_out/bench47_n1.json") if l.startswith("{")][-1])
print("value", d["value"], "ms/step", d["ms_per_step"], "frac", d["roofline"]["frac"], "other", d["roofline"]["other_kernels_ms_per_step"], "e2e", d["e2e"]["seconds_per_step"])
PY
