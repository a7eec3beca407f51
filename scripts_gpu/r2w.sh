#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python tools/bench_reference_gpu.py 10000 3 2>&1 | grep -v Warn | tail -5 > gpurun_out/r2w_refgpu.log
cat gpurun_out/r2w_refgpu.log | cut -c1-800
