"""One line per kernel of an `ncu --set full` capture (`ncu -i X.ncu-rep --page raw --csv > raw.csv`): duration, DRAM
bytes, achieved GB/s against the measured copy peak (MEASURED_PEAKS.json), tensor-pipe and issue utilisation, launch
geometry.  `python tools/summarize_ncu_table.py raw.csv ["header text"] > profiles/<name>.txt`"""
import csv, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
peak = 6554.6
try:
    peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass
rows = list(csv.reader(open(sys.argv[1], errors="replace")))
h0 = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
hdr = rows[h0]
idx = {h: i for i, h in enumerate(hdr)}
num = lambda r, k: float(r[idx[k]].replace(",", "")) if k in idx and r[idx[k]] not in ("", "n/a") else float("nan")
if len(sys.argv) > 2:
    print("\n".join("# " + l for l in sys.argv[2].split("\n")))
print(f"# achieved = (DRAM read + written) / duration against the measured copy peak ({peak:.1f} GB/s); ncu times are cold-cache and serialised")
print(f"# {'kernel':76s} {'us':>7s} {'read MB':>8s} {'write MB':>8s} {'GB/s':>7s} {'of peak':>7s} {'tensor%':>7s} {'issue%':>6s} {'grid':>5s} {'block':>5s} {'regs':>4s}")
units = rows[h0 + 1]
for r in rows[h0 + 2:]:
    if len(r) < len(hdr):
        continue
    t = num(r, "gpu__time_duration.sum")
    tu = units[idx["gpu__time_duration.sum"]]
    t_us = t / 1e3 if tu in ("ns", "nsecond") else (t if tu in ("us", "usecond") else t * 1e3)
    scale = lambda k: {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(units[idx[k]], 1e-6)
    rd = num(r, "dram__bytes_read.sum") * scale("dram__bytes_read.sum")
    wr = num(r, "dram__bytes_write.sum") * scale("dram__bytes_write.sum")
    gbs = (rd + wr) / t_us * 1e3 if t_us == t_us and t_us > 0 else float("nan")
    print(f"  {r[idx['Kernel Name']][:76]:76s} {t_us:7.1f} {rd:8.1f} {wr:8.1f} {gbs:7.0f} {gbs / peak:7.2f} "
          f"{num(r, 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active'):7.1f} "
          f"{num(r, 'smsp__issue_active.avg.pct_of_peak_sustained_active'):6.1f} {num(r, 'launch__grid_size'):5.0f} "
          f"{num(r, 'launch__block_size'):5.0f} {num(r, 'launch__registers_per_thread'):4.0f}")
