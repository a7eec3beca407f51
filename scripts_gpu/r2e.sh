#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
MODALTUNE_B200_LIB=build_exp/libmt_trace.so timeout 120 python tools/attn_trace.py 10001 1024 1 2 > gpurun_out/r2e_trace_1024.log 2>&1
MODALTUNE_B200_LIB=build_exp/libmt_trace.so timeout 120 python tools/attn_trace.py 10001 0 1 2 > gpurun_out/r2e_trace_all.log 2>&1
wc -l gpurun_out/r2e_trace_1024.log gpurun_out/r2e_trace_all.log
