#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
( time timeout 1800 python -m pytest tests -q -m gpu -x ) 2>&1 | grep -v Warning | tail -8 > gpurun_out/r3y_pytest.log
tail -6 gpurun_out/r3y_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | grep smoke
( time timeout 1200 python bench.py ) > gpurun_out/r3y_bench.log 2>&1
python - <<'PY'
import json
t=open("gpurun_out/r3y_bench.log").read()
l=[x for x in t.splitlines() if x.startswith('{')]
print(t[-200:].replace("\n"," | "))
if l:
    d=json.loads(l[-1])
    print("value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"], "steps", d["steps"], "launches", d.get("gpu_launches_per_step"))
    print("roofline", {k: d["roofline"][k] for k in ("frac","achieved","launch_ms","traffic")}, d["roofline"]["fwd"]["frac"])
    print("cpu", d.get("cpu_baseline"))
    print("ref gpu", d.get("reference_gpu_path"))
    print("clocks", d.get("clocks"))
PY
