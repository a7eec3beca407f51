#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py -q -m gpu -x -k "merge or dilated" 2>&1 | grep -v Warning | tail -4
python tools/run_merge_kernels.py 10001 2>&1 | tail -1
python tools/run_merge_kernels.py 32769 2>&1 | tail -1
