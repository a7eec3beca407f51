#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__warps_active.avg.pct_of_peak_sustained_active,launch__grid_size,launch__block_size --clock-control none --csv --log-file gpurun_out/r3b_cross_ncu.csv python tools/run_cross_kernels.py --plain > gpurun_out/r3b_ncu.log 2>&1
python - <<'PY'
import csv
rows=list(csv.reader(l for l in open("gpurun_out/r3b_cross_ncu.csv") if l.startswith('"')))
h=rows[0]; ki=h.index("Kernel Name"); mi=h.index("Metric Name"); vi=h.index("Metric Value"); ii=h.index("ID")
d={}
for r in rows[1:]:
    d.setdefault((int(r[ii]), r[ki][:40]), {})[r[mi]]=r[vi]
for k in sorted(d):
    m=d[k]
    if any(x in k[1] for x in ("cross","Memset","memset")):
        print(k[0], k[1], m.get("gpu__time_duration.sum"), "grid", m.get("launch__grid_size"), "blk", m.get("launch__block_size"), "rd", m.get("dram__bytes_read.sum"), "wr", m.get("dram__bytes_write.sum"), "occ", m.get("sm__warps_active.avg.pct_of_peak_sustained_active"))
PY
