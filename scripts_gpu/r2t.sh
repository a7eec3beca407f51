#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -q -m gpu -x -k "linear_sm100" 2>&1 | tail -30 > gpurun_out/r2t_pytest.log
timeout 300 python tools/bench_linear.py 10001 2>&1 | grep -v Warn > gpurun_out/r2t_bench_linear.log
tail -8 gpurun_out/r2t_pytest.log; cat gpurun_out/r2t_bench_linear.log
