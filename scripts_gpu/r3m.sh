#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r3m_bench.log 2>&1
python - <<PY
import json
t=open("gpurun_out/r3m_bench.log").read()
l=[x for x in t.splitlines() if x.startswith('{')]
if l:
    d=json.loads(l[-1])
    print("value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"], "eager", d.get("eager_ms_per_step"), "launches", d.get("gpu_launches_per_step"), "roof", d["roofline"]["frac"], d["roofline"]["fwd"]["frac"], "loss", d["loss"])
else: print(t[-3000:])
PY
