#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
python tools/bench_train_mode.py 2>&1 | tail -1
MODALTUNE_B200_SPLIT_GRADS=0 python tools/bench_train_mode.py 2>&1 | tail -1
MODALTUNE_B200_INJECTOR_FUSED=0 python tools/bench_train_mode.py 2>&1 | tail -1
MODALTUNE_B200_SHARED_KV=0 python tools/bench_train_mode.py 2>&1 | tail -1
MODALTUNE_B200_CROSS_TC=0 python tools/bench_train_mode.py 2>&1 | tail -1
