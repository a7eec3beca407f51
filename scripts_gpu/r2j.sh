#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
MODALTUNE_B200_LIB=build_exp/libmt_timeline.so timeout 200 python tools/attn_timeline.py 10001 1 2 > gpurun_out/r2j_timeline_10k.log 2>&1
MODALTUNE_B200_LIB=build_exp/libmt_timeline.so timeout 200 python tools/attn_timeline.py 32769 1 2 > gpurun_out/r2j_timeline_32k.log 2>&1
grep -A14 "^persistent backward" gpurun_out/r2j_timeline_10k.log; grep -A14 "^persistent backward" gpurun_out/r2j_timeline_32k.log
