"""Merge the MT_DEBUG_TRACE stamps printed by an attention kernel (block 0) into one timeline."""
import re, sys
ev = []
for line in open(sys.argv[1]):
    m = re.match(r'(\w+) (\d+) (\d+)', line)
    if m and not line.startswith("N="): ev.append((int(m.group(3)), m.group(1), int(m.group(2))))
ev.sort()
lo, hi = (int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (0, 120)
t0 = ev[0][0]; last = {}
for n, (t, w, tag) in enumerate(ev):
    d = t - last.get(w, t); last[w] = t
    if lo <= n < hi: print(f"{t-t0:8d} {w} {tag:5d}  (+{d})")
