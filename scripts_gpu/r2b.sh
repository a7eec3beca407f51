#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -q -m gpu -k "tcgen05_forward" 2>&1 | tail -30 > gpurun_out/r2b_pytest_fwd.log
for n in 10001 32769; do
  timeout 120 python tools/run_attn_kernels.py $n 6 1 1 >> gpurun_out/r2b_attn.log 2>&1
  timeout 120 python tools/run_attn_kernels.py $n 6 1 2 >> gpurun_out/r2b_attn.log 2>&1
done
MODALTUNE_B200_LIB=build_exp/libmt_timeline.so timeout 200 python tools/attn_timeline.py 10001 2 1 > gpurun_out/r2b_timeline_10k.log 2>&1
timeout 900 python -m pytest tests/test_gpu_model.py -q -m gpu -k "encoder_layer_golden or optimizer or full_gradient" 2>&1 | tail -30 > gpurun_out/r2b_pytest_model.log
tail -3 gpurun_out/r2b_pytest_fwd.log; cat gpurun_out/r2b_attn.log; tail -3 gpurun_out/r2b_pytest_model.log
