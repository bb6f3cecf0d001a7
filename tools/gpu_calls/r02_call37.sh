#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "diag or apply or mcfs or median or pooled or statdist or cmstats" > gpurun_out/t37_diag.log 2>&1; echo "diag tests rc=$?"; tail -1 gpurun_out/t37_diag.log
for v in 0 1 2; do
  GCX_DIAG_APPLY_VARIANT=$v timeout 300 python bench.py --workload wg_25kb_mcfs --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/bench37_mcfs_v$v.json 2> gpurun_out/bench37_mcfs_v$v.err; echo "variant $v rc=$? $(python -c "import json;d=json.loads(open('gpurun_out/bench37_mcfs_v$v.json').read().strip().splitlines()[-1]);print(d['ms_per_step'])")"
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_diag_apply_b" -c 1 -o gpurun_out/ncu37_mcfs -f python bench.py --workload wg_25kb_mcfs --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/ncu37.log 2>&1; echo "ncu rc=$?"
