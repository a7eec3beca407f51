#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_model.py -q -m gpu -x -k "ffn_backward_fused or encoder_layer" 2>&1 | grep -v Warning | tail -30 > gpurun_out/r3g_pytest.log
tail -12 gpurun_out/r3g_pytest.log
python tools/run_layer_kernels.py 10001 2 sm100 2>&1 | tail -2
MODALTUNE_B200_FFN_BWD_FUSED=0 python tools/run_layer_kernels.py 10001 2 sm100 2>&1 | tail -1
