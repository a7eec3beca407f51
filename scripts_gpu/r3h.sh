#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
for f in 1 0; do
MODALTUNE_B200_FFN_BWD_FUSED=$f timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r3h_bench_$f.log 2>&1
python - <<PY
import json
t=open("gpurun_out/r3h_bench_$f.log").read()
l=[x for x in t.splitlines() if x.startswith('{')]
if l:
    d=json.loads(l[-1])
    print("fused=$f value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], "eager", d.get("eager_ms_per_step"), "launches", d.get("gpu_launches_per_step"), "roof", d["roofline"]["frac"], d["roofline"]["fwd"]["frac"], "loss", d["loss"])
else: print(t[-2000:])
PY
done
