#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 600 python tools/profile_step.py 10000 --ops 2>&1 | grep -v Warn > gpurun_out/r3f_profile_step.log
head -75 gpurun_out/r3f_profile_step.log | cut -c1-150
