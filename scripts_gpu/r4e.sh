#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_model.py -q -m gpu -x 2>&1 | grep -v Warning | tail -3 | cut -c1-300
for v in 1 0 1 0; do
MODALTUNE_B200_SHARED_KV=$v timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extras 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('shared_kv=$v', 'ms', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], 'roof', d['roofline']['frac'], d['roofline']['fwd']['frac'], 'loss', d['loss'])
"
done
