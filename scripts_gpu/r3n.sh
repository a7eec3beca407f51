#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node ${NG:-2} --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus ${NG:-2} --steps 10 --warmup 3 --no-cpu-baseline ) > gpurun_out/r3n_bench2.log 2>&1
python - <<PY
import json
t=open("gpurun_out/r3n_bench2.log").read()
l=[x for x in t.splitlines() if x.startswith('{')]
print(t[-300:].replace("\n"," | "))
if l:
    d=json.loads(l[-1])
    print("value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], "n", d["n_gpus"])
    print(json.dumps(d.get("extras", {}).get("c5_variable_tiles"))[:1500])
PY
