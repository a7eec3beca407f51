"""Top stall-sample instructions of one kernel from `ncu -i X.ncu-rep --page source --csv --kernel-name regex:K` output."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
h0 = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[h0]; idx = {h: i for i, h in enumerate(hdr)}
data = []
for r in rows[h0 + 1:]:
    if len(r) < len(hdr):
        if r and r[0] == "Kernel Name": break   # next kernel instance
        continue
    data.append(r)
tot = sum(int(r[idx['# Samples']]) for r in data)
print('total samples', tot, 'instructions', len(data))
stalls = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
agg = {h: sum(int(r[idx[h]]) for r in data) for h in stalls}
print('by reason:', sorted(((v, k) for k, v in agg.items()), reverse=True)[:8])
order = sorted(range(len(data)), key=lambda i: -int(data[i][idx['# Samples']]))[:n]
for pos in order:
    r = data[pos]; s = int(r[idx['# Samples']])
    st = sorted(((int(r[idx[h]]), h[6:]) for h in stalls), reverse=True)[:2]
    print(f"{pos:5d} {s:6d} {100*s/tot:5.1f}% {r[idx['Instructions Executed']]:>9} {r[idx['Source']].strip()[:72]:72s} {st}")
