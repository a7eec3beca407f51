#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_model.py tests/test_gpu_kernels.py -q -m gpu -x -k "ffn_backward_fused or encoder_layer or linear" 2>&1 | grep -v Warning | tail -30 > gpurun_out/r3g_pytest.log
tail -5 gpurun_out/r3g_pytest.log
bash scripts_gpu/r3h.sh
