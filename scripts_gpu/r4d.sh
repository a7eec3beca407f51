#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -q -m gpu -x -k "tcgen05_forward" 2>&1 | grep -v Warning | tail -2
for r in 1 2; do
echo packed; python tools/run_attn_kernels.py 10001 20 2 3 2>&1 | tail -1; python tools/run_attn_kernels.py 32769 10 2 3 2>&1 | tail -1
echo unpacked; MODALTUNE_B200_LIB=$PWD/build_exp/lib_k48_unpacked.so python tools/run_attn_kernels.py 10001 20 2 3 2>&1 | tail -1; MODALTUNE_B200_LIB=$PWD/build_exp/lib_k48_unpacked.so python tools/run_attn_kernels.py 32769 10 2 3 2>&1 | tail -1
done
