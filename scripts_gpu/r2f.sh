#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
rm -f gpurun_out/r2g_*.log
for n in 10001 32769; do
timeout 120 python tools/run_attn_kernels.py $n 6 2 1 >> gpurun_out/r2g_attn.log 2>&1 || echo "bwd impl 2 run failed" >> gpurun_out/r2g_attn.log
timeout 120 python tools/run_attn_kernels.py $n 6 1 1 >> gpurun_out/r2g_attn.log 2>&1
done
timeout 900 python -m pytest tests/test_gpu_kernels.py -q -m gpu -x -k "tcgen05_backward" 2>&1 | tail -5 > gpurun_out/r2g_pytest_bwd.log
MODALTUNE_B200_LIB=build_exp/libmt_trace.so timeout 120 python tools/attn_trace.py 10001 1024 1 2 > gpurun_out/r2g_trace_1024.log 2>&1
grep -v Warn gpurun_out/r2g_attn.log; tail -3 gpurun_out/r2g_pytest_bwd.log
