#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -k "linear_sm100 or graphed or train_mode or fused_gemm" 2>&1 | tail -40 > gpurun_out/r2p_pytest.log
tail -6 gpurun_out/r2p_pytest.log
