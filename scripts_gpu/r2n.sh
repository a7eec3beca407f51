#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -q -m gpu -k "linear_sm100" 2>&1 | tail -40 > gpurun_out/r2n_pytest.log
timeout 300 python tools/bench_linear.py 10001 > gpurun_out/r2n_bench_linear.log 2>&1
tail -25 gpurun_out/r2n_pytest.log; grep -v Warn gpurun_out/r2n_bench_linear.log
