#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -40 > gpurun_out/r2o_pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r2o_bench.log 2>&1
MODALTUNE_B200_GEMM=cublas timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r2o_bench_cublas.log 2>&1
timeout 300 python __graft_entry__.py smoke > gpurun_out/r2o_smoke.log 2>&1
tail -4 gpurun_out/r2o_pytest.log; grep smoke gpurun_out/r2o_smoke.log | grep -v print
for f in r2o_bench r2o_bench_cublas; do python - <<PY
import json
l=[x for x in open('gpurun_out/$f.log') if x.startswith('{')]
if l:
    d=json.loads(l[-1]); r=d['roofline']
    print('$f', {k:d[k] for k in ('value','ms_per_step','eager_ms_per_step','gpu_launches_per_step','loss')}, d['e2e']['value'], round(r['frac'],3), round(r['fwd']['frac'],3))
else:
    print('$f: no json'); print(open('gpurun_out/$f.log').read()[-1500:])
PY
done
