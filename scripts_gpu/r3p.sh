#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
L="python tools/run_layer_kernels.py 10001 2 sm100"
X="python tools/run_cross_kernels.py --plain"
B="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-extras"
$L > gpurun_out/r3p_layer_plain.log 2>&1 && \
ncu --set full --clock-control none -k regex:"ln_|linear_sm100|cast_kernel|ffn_bwd_prep|merge_ln" -s 24 -c 24 -f -o gpurun_out/r3_prof_layer $L > gpurun_out/r3p_ncu_layer.log 2>&1
ncu -i gpurun_out/r3_prof_layer.ncu-rep --page raw --csv > gpurun_out/r3_prof_layer_raw.csv 2>/dev/null
$X > gpurun_out/r3p_cross_plain.log 2>&1 && \
ncu --set full --clock-control none -k regex:"cross_.*tc|cross_combine|colsum" -s 0 -c 12 -f -o gpurun_out/r3_prof_cross $X > gpurun_out/r3p_ncu_cross.log 2>&1
ncu -i gpurun_out/r3_prof_cross.ncu-rep --page raw --csv > gpurun_out/r3_prof_cross_raw.csv 2>/dev/null
$B > gpurun_out/r3p_bench_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -s 10500 -c 3200 --csv --log-file gpurun_out/r3_launches.csv $B > gpurun_out/r3p_ncu_bench.log 2>&1
rm -f gpurun_out/r3_prof_layer.ncu-rep gpurun_out/r3_prof_cross.ncu-rep
ls -la gpurun_out/r3_*; tail -2 gpurun_out/r3p_bench_plain.log | cut -c1-400
