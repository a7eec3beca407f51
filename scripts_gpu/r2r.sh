#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
for lib in noepi notma; do
MODALTUNE_B200_LIB=build_exp/libmt_$lib.so timeout 300 python tools/bench_linear.py 10001 2>&1 | grep -v Warn > gpurun_out/r2r_linear_$lib.log
done
