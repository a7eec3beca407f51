#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_model.py -q -m gpu -x -k "graphed_step_drives" 2>&1 | grep -v Warning | head -60 > gpurun_out/r3e_pytest.log
head -50 gpurun_out/r3e_pytest.log
