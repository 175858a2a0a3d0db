"""Per-kernel totals of an ncu launch list of one training step (tools/gpu_train_round.sh -> gpurun_out/train_launches.csv).
    python tools/train_launch_summary.py [csv] [--top N]   # N slowest single launches as well"""
import collections
import csv
import re
import sys

path = next((a for a in sys.argv[1:] if not a.startswith("--")), "gpurun_out/train_launches.csv")
top = int(sys.argv[sys.argv.index("--top") + 1]) if "--top" in sys.argv else 0
with open(path) as f:
    lines = [l for l in f if not l.startswith("==")]
agg = collections.defaultdict(lambda: [0, 0.0])
tot, single = 0.0, []
for row in csv.DictReader(lines):
    if row.get("Metric Name") != "gpu__time_duration.sum":
        continue
    name, v = row["Kernel Name"], float(row["Metric Value"].replace(",", ""))
    v = v / 1e3 if row["Metric Unit"] == "ns" else v * 1e3 if row["Metric Unit"] == "ms" else v
    m = re.search(r"ts_(?:row_)?kernel<.*?::(\w+)>", name)
    t = re.search(r"ts_bmm_tiled<(\d+), (\d+)", name)
    key = m.group(1) if m else f"bmm {t.group(1)}x{t.group(2)}" if t else name.split("(")[0][:48]
    agg[key][0] += 1
    agg[key][1] += v
    tot += v
    single.append((v, key, row.get("Grid Size", ""), row["ID"]))
for k, (n, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
    print(f"{k:40s} {n:5d} launches {t / 1e3:9.3f} ms {100 * t / tot:5.1f} %")
print(f"total {tot / 1e3:.3f} ms (cold cache, serialised)")
for v, k, g, i in sorted(single, reverse=True)[:top]:
    print(f"  {v:9.1f} us  {k:24s} grid {g} id {i}")
