"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count, total time, share."""
import csv, sys, collections, re
path = sys.argv[1]
rows = [r for r in csv.reader(open(path, errors="replace")) if len(r) > 5]
hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
cols = {n: i for i, n in enumerate(rows[hdr])}
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows[hdr + 1:]:
    try:
        name, val, unit = r[cols["Kernel Name"]], float(r[cols["Metric Value"]].replace(",", "")), r[cols["Metric Unit"]]
    except Exception:
        continue
    ns = val * {"ns": 1, "us": 1e3, "ms": 1e6, "s": 1e9}.get(unit.replace("second", "s"), 1)
    name = re.sub(r"\(.*", "", name)[:90]
    agg[name][0] += 1
    agg[name][1] += ns
tot = sum(v[1] for v in agg.values())
print(f"# {path}: {sum(v[0] for v in agg.values())} launches, {tot/1e6:.2f} ms total device time (serialised, cold cache)")
print(f"{'ms':>9} {'count':>6} {'share':>6}  kernel")
for name, (c, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:40]:
    print(f"{ns/1e6:9.3f} {c:6d} {100*ns/tot:5.1f}%  {name}")
