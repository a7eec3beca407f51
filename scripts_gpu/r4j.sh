#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
timeout 1500 python -m pytest tests/test_gpu_kernels.py -q -m gpu -x -k "dilated or merge" 2>&1 | grep -v Warning | tail -3
timeout 300 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-extras 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('ms', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], d['roofline']['fwd']['kernel'])
"
