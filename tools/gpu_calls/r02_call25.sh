#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
python tools/profile_ic.py 24926 > gpurun_out/plain25.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_ic_pass_symd -c 1 -o gpurun_out/r02_ic_pass_symd_final python tools/profile_ic.py 24926 > gpurun_out/ncu25.log 2>&1; echo "ncu full rc=$?"
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench25_ref.json 2> gpurun_out/bench25_ref.err; echo "ref rc=$?"
timeout 300 python bench.py --workload chr1_10kb_corr --steps 3 --warmup 1 > gpurun_out/corr25_int.json 2> /dev/null; echo "corr rc=$?"
timeout 300 python bench.py --workload chr1_10kb_corr --corr-float --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/corr25_float.json 2> /dev/null
timeout 300 python bench.py --workload chr1_10kb_statdist --steps 3 --warmup 1 --no-cpu-baseline > gpurun_out/statdist25.json 2> /dev/null; echo "sd rc=$?"
timeout 300 python bench.py --workload chr21_100kb --steps 5 --warmup 2 > gpurun_out/chr21_25.json 2> /dev/null; echo "chr21 rc=$?"
timeout 300 python bench.py --workload wg_25kb_mcfs --steps 5 --warmup 2 --no-cpu-baseline > gpurun_out/mcfs25.json 2> /dev/null; echo "mcfs rc=$?"
python -c "
import __graft_entry__ as g
g.smoke()
"
