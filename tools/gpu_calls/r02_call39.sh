#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
GCX_DIAG_APPLY_VARIANT=3 timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "mcfs or early_exit" > gpurun_out/t39_v3.log 2>&1; echo "v3 tests rc=$?"; tail -1 gpurun_out/t39_v3.log
GCX_DIAG_APPLY_VARIANT=4 timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "mcfs or early_exit" > gpurun_out/t39_v4.log 2>&1; echo "v4 tests rc=$?"; tail -1 gpurun_out/t39_v4.log
for v in 3 4; do
  GCX_DIAG_APPLY_VARIANT=$v timeout 300 python bench.py --workload wg_25kb_mcfs --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/bench39_mcfs_v$v.json 2> gpurun_out/bench39_mcfs_v$v.err; echo "variant $v rc=$? $(python -c "import json;d=json.loads(open('gpurun_out/bench39_mcfs_v$v.json').read().strip().splitlines()[-1]);print(d['ms_per_step'])")"
done
GCX_DIAG_APPLY_VARIANT=3 timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_diag_apply" -c 1 -o gpurun_out/ncu39_mcfs -f python bench.py --workload wg_25kb_mcfs --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/ncu39.log 2>&1; echo "ncu rc=$?"
