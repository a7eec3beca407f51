#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_model.py -q -m gpu -x -k "graph_cache or graphed" 2>&1 | grep -v Warning | tail -40 > gpurun_out/r3k_pytest.log
tail -25 gpurun_out/r3k_pytest.log | cut -c1-400
( time timeout 1200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline ) > gpurun_out/r3k_bench.log 2>&1
python - <<'PY'
import json
t=open("gpurun_out/r3k_bench.log").read()
l=[x for x in t.splitlines() if x.startswith('{')]
print(t[-200:].replace("\n"," | "))
if l:
    d=json.loads(l[-1])
    print("value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"])
    print(json.dumps(d.get("extras"))[:3000])
PY
