#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
( time timeout 1200 python bench.py --steps 10 --warmup 3 ) > gpurun_out/r2x_bench_full.log 2>&1
( time timeout 600 python bench.py --impl reference --steps 20 --warmup 3 ) > gpurun_out/r2x_bench_ref.log 2>&1
python - <<'PY'
import json
for f in ("gpurun_out/r2x_bench_full.log", "gpurun_out/r2x_bench_ref.log"):
    t=open(f).read()
    l=[x for x in t.splitlines() if x.startswith('{')]
    print(f, t[-200:].replace("\n"," | "))
    if l:
        d=json.loads(l[-1])
        for k in ("value","ms_per_step","steps","warmup","e2e","extras","reference_gpu_path","cpu_baseline"):
            if k in d: print("  ",k, json.dumps(d[k])[:900])
        if "fa2_comparator" in d: print("   fa2", json.dumps({k:d["fa2_comparator"].get(k) for k in ("speedup_vs_fa2_kernels","speedup_vs_fa2_with_gather_scatter")}))
PY
