#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
timeout 900 python -m pytest tests/test_gpu_model.py -q -m gpu -x -k "encoder_layer or graph or step" 2>&1 | grep -v Warning | tail -2
for v in 1 0 1 0; do
MODALTUNE_B200_ZERO_EARLY=$v timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extras 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('zero_early=$v', 'ms', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], 'loss', d['loss'])
"
done
