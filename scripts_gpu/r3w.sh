#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_model.py -q -m gpu -x -k "graph" 2>&1 | grep -v Warning | tail -4
NG=2 bash scripts_gpu/r3n.sh
