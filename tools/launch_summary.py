"""Summarise an ncu gpu__time_duration launch list (csv): per-kernel totals and the launch sequence of one step."""
import collections, csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
h = rows[hdr]
ki, vi, gi = h.index("Kernel Name"), h.index("Metric Value"), h.index("Grid Size")
agg = collections.OrderedDict(); tot = 0.0; seq = []
for r in rows[hdr + 2:]:
    if len(r) <= vi: continue
    t = float(r[vi].replace(",", "")) / 1000
    name = r[ki].split("(")[0].replace("dhg::", "").replace("<unnamed>::", "")[:60]
    a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += t; tot += t
    seq.append((t, name, r[gi]))
print(f"total {tot:.0f} us over {len(seq)} launches")
for k, (n, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
    print(f"{t:9.1f} us {100*t/tot:5.1f}%  x{n:3d}  {k}")
if len(sys.argv) > 2:
    for t, name, g in seq: print(f"{t:9.1f}  {name[:50]:50s} {g}")
