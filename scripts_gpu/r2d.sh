#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 120 python tools/run_attn_kernels.py 10001 6 2 1 > gpurun_out/r2d_attn.log 2>&1 || echo "bwd impl 2 run failed" >> gpurun_out/r2d_attn.log
timeout 120 python tools/run_attn_kernels.py 10001 6 1 1 >> gpurun_out/r2d_attn.log 2>&1
timeout 120 python tools/run_attn_kernels.py 32769 6 2 1 >> gpurun_out/r2d_attn.log 2>&1
timeout 120 python tools/run_attn_kernels.py 32769 6 1 1 >> gpurun_out/r2d_attn.log 2>&1
timeout 900 python -m pytest tests/test_gpu_kernels.py -q -m gpu -x -k "tcgen05_backward" 2>&1 | tail -30 > gpurun_out/r2d_pytest_bwd.log
MODALTUNE_B200_LIB=build_exp/libmt_timeline.so timeout 200 python tools/attn_timeline.py 10001 1 2 > gpurun_out/r2d_timeline_10k.log 2>&1
timeout 300 python tools/attn_branch_sweep.py 10001 1 2 > gpurun_out/r2d_sweep_bwd2.log 2>&1
grep -v Warn gpurun_out/r2d_attn.log; tail -5 gpurun_out/r2d_pytest_bwd.log; grep -A3 "^backward" gpurun_out/r2d_timeline_10k.log; tail -4 gpurun_out/r2d_sweep_bwd2.log
