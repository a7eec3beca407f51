#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out; rm -f gpurun_out/r2k_*.log
for n in 10001 32769; do
timeout 120 python tools/run_attn_kernels.py $n 6 2 1 >> gpurun_out/r2k_attn.log 2>&1 || echo "bwd impl 2 run failed" >> gpurun_out/r2k_attn.log
timeout 120 python tools/run_attn_kernels.py $n 6 1 1 >> gpurun_out/r2k_attn.log 2>&1
done
timeout 900 python -m pytest tests/test_gpu_kernels.py -q -m gpu -x -k "tcgen05_backward" 2>&1 | tail -5 > gpurun_out/r2k_pytest_bwd.log
MODALTUNE_B200_LIB=build_exp/libmt_timeline.so timeout 200 python tools/attn_timeline.py 10001 1 2 > gpurun_out/r2k_timeline_10k.log 2>&1
grep -v Warn gpurun_out/r2k_attn.log; tail -3 gpurun_out/r2k_pytest_bwd.log; grep -A3 "^backward" gpurun_out/r2k_timeline_10k.log; grep -A10 "^persistent backward" gpurun_out/r2k_timeline_10k.log
