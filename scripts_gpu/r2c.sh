#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 300 python tools/attn_branch_sweep.py 10001 2 1 > gpurun_out/r2c_sweep_fwd2.log 2>&1
grep -v Warn gpurun_out/r2c_sweep_fwd2.log
