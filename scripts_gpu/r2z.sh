#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
A="python tools/run_attn_kernels.py 10001 3 2 1"
L="python tools/run_layer_kernels.py 10001 2 sm100"
B="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-extras"
$A > gpurun_out/r2z_attn_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:dilated_ -s 4 -c 2 -f -o gpurun_out/r2_prof_attn $A > gpurun_out/r2z_ncu_attn.log 2>&1
$L > gpurun_out/r2z_layer_plain.log 2>&1 && \
ncu --set full --clock-control none -k regex:"ln_|linear_sm100|cast_kernel|residual" -s 20 -c 20 -f -o gpurun_out/r2_prof_layer $L > gpurun_out/r2z_ncu_layer.log 2>&1
$B > gpurun_out/r2z_bench_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -s 11000 -c 3600 --csv --log-file gpurun_out/r2_launches.csv $B > gpurun_out/r2z_ncu_bench.log 2>&1
ls -la gpurun_out/r2_prof_attn.ncu-rep gpurun_out/r2_prof_layer.ncu-rep gpurun_out/r2_launches.csv; grep -v Warn gpurun_out/r2z_attn_plain.log gpurun_out/r2z_layer_plain.log | tail -4
