#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -q -m gpu -x -k "cross_attention" 2>&1 | tail -5 > gpurun_out/r3a_pytest.log
timeout 300 python tools/run_cross_kernels.py > gpurun_out/r3a_cross.log 2>&1
cat gpurun_out/r3a_pytest.log; cat gpurun_out/r3a_cross.log
bash scripts_gpu/r3b.sh | grep -v "_kernel<"
