#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_model.py -q -m gpu -x -k "linear or encoder_layer or ffn_backward" 2>&1 | grep -v Warning | tail -5
timeout 300 python tools/bench_linear.py 2>&1 | tail -14
bash scripts_gpu/r3m.sh
