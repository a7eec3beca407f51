#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_train_mode.py -q -m gpu -x 2>&1 | grep -v Warning | tail -2
timeout 900 python bench.py --steps 6 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('ms', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step']); print(json.dumps(d['extras']['train_mode']))
"
