#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/r3i_layer_ncu.csv python tools/run_layer_kernels.py 10001 2 sm100 > gpurun_out/r3i_ncu.log 2>&1
python - <<'PY'
import csv
rows=list(csv.reader(l for l in open("gpurun_out/r3i_layer_ncu.csv") if l.startswith('"')))
h=rows[0]; ki=h.index("Kernel Name"); mi=h.index("Metric Name"); vi=h.index("Metric Value"); ii=h.index("ID")
d={}
for r in rows[1:]:
    d.setdefault((int(r[ii]), r[ki][:44]), {})[r[mi]]=r[vi]
ks=sorted(d)
# last layer pass only: find the last dilated_fwd
last=[k for k in ks if "dilated_fwd" in k[1]][-1][0]
for k in ks:
    if k[0] >= last-6:
        m=d[k]
        print(k[0], k[1], m.get("gpu__time_duration.sum"), "rd", m.get("dram__bytes_read.sum"), "wr", m.get("dram__bytes_write.sum"))
PY
