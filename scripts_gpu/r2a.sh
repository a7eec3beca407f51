#!/bin/bash
# round 2, first GPU call: parity on the shipped path, diagnostics, comparator, baseline bench
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r2a_smi.log 2>&1
free -g > gpurun_out/r2a_free.log; nproc >> gpurun_out/r2a_free.log
timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -60 > gpurun_out/r2a_pytest.log
timeout 600 python tools/diag_grad_cosine.py 520 > gpurun_out/r2a_diag_cos.log 2>&1
timeout 300 python tools/bench_fa2_branches.py 10001 32769 > gpurun_out/r2a_fa2.log 2>&1
timeout 300 python tools/attn_branch_sweep.py 10001 > gpurun_out/r2a_sweep_10k.log 2>&1
MODALTUNE_B200_LIB=build_exp/libmt_timeline.so timeout 200 python tools/attn_timeline.py 10001 > gpurun_out/r2a_timeline_10k.log 2>&1
MODALTUNE_B200_LIB=build_exp/libmt_timeline.so timeout 200 python tools/attn_timeline.py 32769 > gpurun_out/r2a_timeline_32k.log 2>&1
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r2a_bench.log 2>&1
timeout 300 python __graft_entry__.py smoke > gpurun_out/r2a_smoke.log 2>&1
tail -3 gpurun_out/r2a_pytest.log
