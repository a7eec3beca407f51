#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
python tools/bench_train_mode.py 2>&1 | tail -1
timeout 900 python -m pytest tests/test_gpu_model.py tests/test_gpu_train_mode.py -q -m gpu -x -k "graph or train or optimizer" 2>&1 | grep -v Warning | tail -2
