#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out; rm -f gpurun_out/r2i_*.log
for lib in modaltune_b200/libmodaltune_b200.so build_exp/libmt_skipdkv.so build_exp/libmt_skipall.so; do
  echo "== $lib" >> gpurun_out/r2i_attn.log
  MODALTUNE_B200_LIB=$lib timeout 120 python tools/run_attn_kernels.py 10001 6 2 1 >> gpurun_out/r2i_attn.log 2>&1
  MODALTUNE_B200_LIB=$lib timeout 120 python tools/run_attn_kernels.py 32769 6 2 1 >> gpurun_out/r2i_attn.log 2>&1
done
MODALTUNE_B200_LIB=build_exp/libmt_skipall.so timeout 300 python tools/attn_branch_sweep.py 10001 1 2 > gpurun_out/r2i_sweep_skipall.log 2>&1
timeout 300 python tools/attn_branch_sweep.py 10001 1 2 > gpurun_out/r2i_sweep.log 2>&1
grep -v Warn gpurun_out/r2i_attn.log
