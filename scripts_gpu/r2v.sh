#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
ls oracle/_ref/reference | head -3
timeout 1700 python -m pytest tests/test_launcher.py -q -m gpu -x 2>&1 | tail -40 > gpurun_out/r2v_pytest_launcher.log
tail -25 gpurun_out/r2v_pytest_launcher.log
