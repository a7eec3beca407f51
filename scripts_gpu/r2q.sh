#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 300 python tools/bench_linear.py 10001 2>&1 | grep -v Warn > gpurun_out/r2q_linear.log
MODALTUNE_B200_LIB=build_exp/libmt_halfb.so timeout 300 python tools/bench_linear.py 10001 2>&1 | grep -v Warn > gpurun_out/r2q_linear_halfb.log
cat gpurun_out/r2q_linear.log gpurun_out/r2q_linear_halfb.log | cut -c1-400 | awk -F'|' '{print $1 "|" $2}' | cut -c1-60,150-
