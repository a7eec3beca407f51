#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_train_mode.py tests/test_gpu_model.py -q -m gpu -x 2>&1 | tail -5 > gpurun_out/r2y_pytest.log
( time timeout 1200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline ) > gpurun_out/r2y_bench.log 2>&1
tail -3 gpurun_out/r2y_pytest.log
python - <<'PY'
import json
t=open("gpurun_out/r2y_bench.log").read()
l=[x for x in t.splitlines() if x.startswith('{')]
print(t[-150:].replace("\n"," | "))
if l:
    d=json.loads(l[-1])
    print(d["value"], d["ms_per_step"], d["e2e"]["value"])
    print(json.dumps(d.get("extras"))[:1800])
PY
