"""Summarise an `ncu --set full` report: `ncu -i X.ncu-rep --page raw --csv > raw.csv; python tools/summarize_ncu.py raw.csv`.

Prints, per profiled kernel, the metrics DESIGN.md / bench.py quote (duration, DRAM bytes, pipe utilisation, issue rate,
occupancy, L1/L2 throughput, launch geometry) and the top warp-stall reasons."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1], errors="replace")))
h0 = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
hdr, units = rows[h0], rows[h0 + 1]
idx = {h: i for i, h in enumerate(hdr)}
WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "l1tex__throughput.avg.pct_of_peak_sustained_active", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__shared_mem_per_block_dynamic", "smsp__inst_executed.sum", "sm__cycles_elapsed.max",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]
for r in rows[h0 + 2:]:
    if len(r) < len(hdr):
        continue
    print(f"kernel: {r[idx['Kernel Name']][:78]}")
    for m in WANT:
        if m in idx:
            print(f"  {m:82s} {r[idx[m]]:>16s} {units[idx[m]]}")
    stalls = []
    for h, i in idx.items():
        if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio"):
            try:
                stalls.append((float(r[i].replace(",", "")), h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]))
            except ValueError:
                pass
    if stalls:
        print("  top warp-stall reasons (warps stalled per issue-active cycle):")
        for v, n in sorted(stalls, reverse=True)[:6]:
            print(f"    {v:7.3f}  {n}")
    print()
