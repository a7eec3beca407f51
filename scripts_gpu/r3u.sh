#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
for v in old new old new; do
  if [ $v = old ]; then export MODALTUNE_B200_LIB=$PWD/build_exp/lib_old.so; else unset MODALTUNE_B200_LIB; fi
  echo "== $v"
  timeout 300 python tools/bench_linear.py 2>&1 | grep -E "FFN forward|fc1|out_proj" | cut -c1-250
  timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extras 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('$v', 'ms', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], 'roof', d['roofline']['frac'])
"
done
