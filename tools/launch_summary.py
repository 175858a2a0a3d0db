"""Summarise an ncu launch list (csv with one row per launch and metric): per-kernel time totals, optional DRAM bytes,
the launch sequence, and (with --json) the DRAM traffic of one denoiser step for bench.py's roofline.traffic.

    python tools/launch_summary.py launches.csv [--seq] [--step-json out.json]
"""
import collections
import csv
import json
import re
import sys

path = sys.argv[1]
rows = list(csv.reader(open(path)))
hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
h = rows[hdr]
ci = {n: i for i, n in enumerate(h)}
launches = collections.OrderedDict()   # id -> dict
for r in rows[hdr + 1:]:
    if len(r) <= ci["Metric Value"] or not r[ci["ID"]].strip().isdigit():
        continue
    d = launches.setdefault(int(r[ci["ID"]]), {"name": r[ci["Kernel Name"]], "grid": r[ci["Grid Size"]]})
    val = float(r[ci["Metric Value"]].replace(",", ""))
    unit = r[ci["Metric Unit"]]
    m = r[ci["Metric Name"]]
    if m.startswith("gpu__time_duration"):
        d["us"] = val / 1e3 if unit.startswith("ns") else val * (1e3 if unit.startswith("ms") else 1.0)
    elif m.startswith("dram__bytes"):
        scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1.0)
        d["rd" if "read" in m else "wr"] = val * scale


def short(n):
    n = re.sub(r"^void ", "", n).replace("dhg::", "").replace("(anonymous namespace)::", "").replace("<unnamed>::", "")
    return n.split("(")[0][:64]


seq = [(short(d["name"]), d.get("us", 0.0), d.get("rd", 0.0), d.get("wr", 0.0), d["grid"]) for d in launches.values()]
tot = sum(s[1] for s in seq)
agg = collections.OrderedDict()
for n, us, rd, wr, _ in seq:
    a = agg.setdefault(n, [0, 0.0, 0.0])
    a[0] += 1; a[1] += us; a[2] += rd + wr
print(f"total {tot:.0f} us over {len(seq)} launches (ncu: cold cache, serialised -- compare shares, not absolutes)")
for k, (n, t, by) in sorted(agg.items(), key=lambda x: -x[1][1]):
    print(f"{t:9.1f} us {100 * t / tot:5.1f}%  x{n:3d}  {by / 1e6:9.1f} MB dram  {k}")
if "--seq" in sys.argv:
    for n, us, rd, wr, g in seq:
        print(f"{us:9.1f}  {rd / 1e6:8.1f} {wr / 1e6:8.1f}  {n[:56]:56s} {g}")
if "--step-json" in sys.argv:
    # one denoiser step = the launches between two consecutive head kernels (the chain runs without a graph)
    heads = [i for i, s in enumerate(seq) if s[0].startswith("heads_")]
    if len(heads) >= 4:
        # the text sides of two consecutive steps are launched together (engine.cu run_chain), so average over a pair of steps
        a, b = heads[1] + 1, heads[3] + 1
        step = seq[a:b]
        out = {
            "launches_per_step": len(step) / 2,
            "step_us_ncu": sum(s[1] for s in step) / 2,
            "step_dram_bytes": sum(s[2] + s[3] for s in step) / 2,
            "step_dram_read_bytes": sum(s[2] for s in step) / 2,
            "step_dram_write_bytes": sum(s[3] for s in step) / 2,
            "gemm_dram_bytes_per_step": sum(s[2] + s[3] for s in step if s[0].startswith("tc_gemm_kernel")) / 2,
            "gemm_launches_per_step": sum(1 for s in step if s[0].startswith("tc_gemm_kernel")) / 2,
            "attention_dram_bytes_per_step": sum(s[2] + s[3] for s in step if s[0].startswith("attn_tc")) / 2,
            "by_kernel_us": {k: round(sum(s[1] for s in step if s[0] == k) / 2, 1) for k in sorted({s[0] for s in step})},
            "source": "ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none, "
                      "tools/profile_chain.py 1024 bf16 (chain without CUDA graph, DHG_OPTS=autotune=0); average of the second and third step of the chain",
        }
        json.dump(out, open(sys.argv[sys.argv.index("--step-json") + 1], "w"), indent=1)
        print("step:", json.dumps({k: v for k, v in out.items() if k != "by_kernel_us" and k != "source"}))
