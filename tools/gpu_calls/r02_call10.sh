#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/t_gpu10.log 2>&1; echo "gpu rc=$?"; tail -4 gpurun_out/t_gpu10.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench10_n1.json 2> gpurun_out/bench10_n1.err; echo "bench rc=$?"
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench10_ref.json 2> gpurun_out/bench10_ref.err; echo "ref rc=$?"
python tools/profile_ic.py 24926 > gpurun_out/plain10.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r02_launches_ic.csv python tools/profile_ic.py 24926 > gpurun_out/ncu10a.log 2>&1; echo "ncu launches rc=$?"
python tools/profile_ic.py 24926 > gpurun_out/plain10.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_ic_pass_symd -c 1 -o gpurun_out/r02_ic_pass_symd_n24926 python tools/profile_ic.py 24926 > gpurun_out/ncu10b.log 2>&1; echo "ncu full rc=$?"
tail -3 gpurun_out/ncu10b.log
