#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_model.py tests/test_gpu_train_mode.py -q -m gpu -x 2>&1 | grep -v Warning | tail -4 | cut -c1-300
bash scripts_gpu/r3m.sh
