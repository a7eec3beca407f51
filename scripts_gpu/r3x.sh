#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
B="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-extras"
ncu --metrics gpu__time_duration.sum --clock-control none -s 10000 -c 2900 --csv --log-file gpurun_out/r3x_launches.csv $B > gpurun_out/r3x_ncu_bench.log 2>&1
python tools/summarize_launches.py gpurun_out/r3x_launches.csv | head -60
