#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -x 2>&1 | grep -v Warning | tail -60 > gpurun_out/r3d_pytest.log
( time timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extras ) > gpurun_out/r3d_bench.log 2>&1
tail -5 gpurun_out/r3d_pytest.log
python - <<'PY'
import json
t=open("gpurun_out/r3d_bench.log").read()
l=[x for x in t.splitlines() if x.startswith('{')]
print(t[-300:].replace("\n"," | "))
if l:
    d=json.loads(l[-1])
    print("value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], "eager", d.get("eager_ms_per_step"), "launches", d.get("gpu_launches_per_step"), "roof", d["roofline"]["frac"], d["roofline"]["fwd"]["frac"], "loss", d["loss"])
PY
