#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 600 python tools/profile_step.py 10000 2>&1 | grep -v Warn > gpurun_out/r2u_profile_step.log
head -60 gpurun_out/r2u_profile_step.log
