#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "corr or own_int8 or correlation" > gpurun_out/t_corr20.log 2>&1; echo "corr tests rc=$?"; tail -3 gpurun_out/t_corr20.log
timeout 300 python bench.py --workload chr1_10kb_corr --steps 3 --warmup 1 --no-cpu-baseline > gpurun_out/corr20_int.json 2> gpurun_out/corr20_int.err; echo "corr int rc=$?"
timeout 300 python bench.py --workload chr1_10kb_corr --corr-float --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/corr20_float.json 2> /dev/null
python - <<'PY'
import json
for f in ("corr20_int.json", "corr20_float.json"):
    try:
        d = json.loads(open("gpurun_out/" + f).read().strip().splitlines()[-1]); print(f, d["value"], d["roofline"]["achieved"])
    except Exception as e: print(f, "unreadable", e)
PY
python tools/profile_corr.py > gpurun_out/plain20.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/r02_launches_corr2.csv python tools/profile_corr.py > gpurun_out/ncu20.log 2>&1; echo "ncu rc=$?"
