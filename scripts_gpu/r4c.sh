#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
A="python tools/run_attn_kernels.py 10001 3 2 3"
$A > gpurun_out/r4c_attn_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:dilated_ -s 4 -c 2 -f -o gpurun_out/r4_prof_attn $A > gpurun_out/r4c_ncu_attn.log 2>&1
ncu -i gpurun_out/r4_prof_attn.ncu-rep --page raw --csv > gpurun_out/r4_prof_attn_raw.csv 2>/dev/null
rm -f gpurun_out/r4_prof_attn.ncu-rep
python tools/summarize_ncu.py gpurun_out/r4_prof_attn_raw.csv | head -80
