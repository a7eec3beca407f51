#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
MODALTUNE_B200_LIB=build_exp/libmt_gtrace.so timeout 120 python tools/gemm_trace.py 2>&1 | grep -v Warn > gpurun_out/r2s_gemm_trace.log
cat gpurun_out/r2s_gemm_trace.log
