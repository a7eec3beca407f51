#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -q -m gpu -x -k "tcgen05_forward" 2>&1 | grep -v Warning | tail -2
echo "default lib impl 1 / 3"
for impl in 1 3; do python tools/run_attn_kernels.py 10001 20 2 $impl 2>&1 | tail -1; done
for f in build_exp/lib_k*_s*.so; do echo $f; MODALTUNE_B200_LIB=$PWD/$f python tools/run_attn_kernels.py 10001 20 2 3 2>&1 | tail -1; MODALTUNE_B200_LIB=$PWD/$f python tools/run_attn_kernels.py 32769 10 2 3 2>&1 | tail -1; done
