#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "deferred or at_size or symmetric_pass or mcfs or fused_apply or ic_vs_fp64" > gpurun_out/t_quick3.log 2>&1; echo "quick rc=$?"; tail -4 gpurun_out/t_quick3.log
timeout 300 python tools/sym_phase_stamps.py 24926 > gpurun_out/stamps2_n24926.txt 2>&1; echo "stamps rc=$?"
cat gpurun_out/stamps2_n24926.txt
GCX_SYM_POOL=0 timeout 300 python tools/sym_phase_stamps.py 24926 > gpurun_out/stamps2_pool0.txt 2>&1
head -12 gpurun_out/stamps2_pool0.txt
timeout 600 python tools/sym_sweep.py 24926 5,32 10,16 10,32 10,64 20,16 20,32 20,64 30,16 30,32 30,64 40,32 > gpurun_out/sym_sweep2.txt 2>&1; echo "sweep rc=$?"
cat gpurun_out/sym_sweep2.txt
timeout 300 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-e2e > gpurun_out/bench3.json 2> gpurun_out/bench3.err; echo "bench rc=$?"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench3.json").read().strip().splitlines()[-1])
print({k: d.get(k) for k in ("value", "ms_per_step", "time_to_converge_s", "passes", "products_launched")})
print(d["roofline"]); print(d.get("mcfs")); print(d.get("kr"))
PY
