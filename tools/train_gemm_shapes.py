"""Joins the contractions of a training-step plan (listed by the host build of csrc/train_step.cu with DHG_TRAINER_LOG_BMM=1,
no arithmetic) with the per-launch times of an ncu launch list of the same step on the GPU: TFLOP/s per launch, worst first.

    python tools/train_gemm_shapes.py [gpurun_out/train_launches.csv] [B T L]
"""
import collections
import csv
import ctypes
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if os.environ.get("DHG_TRAINER_LOG_BMM") == "1" and "--child" in sys.argv:
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    sys.path.insert(0, os.path.join(ROOT, "diffusion-handwriting-generation.pytorch_b200"))
    import torch
    import hostsim_build

    B, T, L = (int(a) for a in sys.argv[sys.argv.index("--child") + 1:][:3])
    h = hostsim_build.lib()
    n = h.dhg_trainer_param_count(2, 128)
    p, g = torch.zeros(n), torch.zeros(n)
    tr = ctypes.c_void_p()
    assert h.dhg_trainer_create(0, 2, 128, B, T, L, p.data_ptr(), g.data_ptr(), ctypes.byref(tr)) == 0
    x, text, sig, style = torch.zeros(B, T, 2), torch.ones(B, L, dtype=torch.int64), torch.ones(B), torch.zeros(B, 14, 1280)
    sys.stdout.flush()
    h.dhg_trainer_forward(tr, x.data_ptr(), text.data_ptr(), sig.data_ptr(), style.data_ptr(), None, None, None, None)
    h.dhg_trainer_backward(tr, x.data_ptr(), torch.zeros(B, T).data_ptr(), None)
    sys.exit(0)

args = [a for a in sys.argv[1:] if not a.startswith("--")]
path = args[0] if args and not args[0].isdigit() else os.path.join(ROOT, "gpurun_out", "train_launches.csv")
B, T, L = (int(a) for a in args[-3:]) if len(args) >= 3 and args[-1].isdigit() else (96, 480, 50)
out = subprocess.run([sys.executable, __file__, "--child", str(B), str(T), str(L)], env=dict(os.environ, DHG_TRAINER_LOG_BMM="1"),
                     stdout=subprocess.PIPE, text=True, check=True).stdout
shapes = [tuple(int(v) for v in l.split()[1:]) for l in out.splitlines() if l.startswith("BMM ")]
with open(path) as f:
    rows = [r for r in csv.DictReader(l for l in f if not l.startswith("==")) if r.get("Metric Name") == "gpu__time_duration.sum"]
times = []
for r in rows:
    if "ts_bmm_tiled" not in r["Kernel Name"]:
        continue
    v = float(r["Metric Value"].replace(",", ""))
    v = v / 1e3 if r["Metric Unit"] == "ns" else v * 1e3 if r["Metric Unit"] == "ms" else v
    t = re.search(r"ts_bmm_tiled<(\d+), (\d+)", r["Kernel Name"])
    times.append((v, f"{t.group(1)}x{t.group(2)}", r.get("Grid Size", "")))
assert len(times) == len(shapes), (len(times), len(shapes))
MODE = {0: "store", 1: "accumulate", 2: "atomic (weight grad)"}
tot_t = sum(t for t, _, _ in times)
tot_f = 0.0
groups = collections.defaultdict(lambda: [0, 0.0, 0.0])
items = []
for (us, tile, grid), (M, N, K, Z, taps, mode) in zip(times, shapes):
    flop = 2.0 * M * N * K * Z * taps
    tot_f += flop
    items.append((us, flop / us / 1e6, M, N, K, Z, taps, mode, tile, grid))
    key = (MODE[mode], "conv" if taps == 3 else "batched (attention)" if Z > 1 and mode != 2 else "flat")
    groups[key][0] += 1
    groups[key][1] += us
    groups[key][2] += flop
print(f"{len(items)} contractions, {tot_f / 1e9:.1f} GFLOP in {tot_t / 1e3:.2f} ms = {tot_f / tot_t / 1e6:.1f} TFLOP/s (ncu: cold cache, serialised)")
for k, (n, us, fl) in sorted(groups.items(), key=lambda kv: -kv[1][1]):
    print(f"  {k[0]:22s} {k[1]:22s} {n:4d} launches {us / 1e3:7.2f} ms {fl / 1e9:7.1f} GFLOP {fl / us / 1e6:6.1f} TFLOP/s")
print("slowest launches:")
for us, tf, M, N, K, Z, taps, mode, tile, grid in sorted(items, reverse=True)[:30]:
    print(f"  {us:7.1f} us {tf:5.1f} TFLOP/s  M={M:<6d} N={N:<5d} K={K:<5d} Z={Z:<4d} taps={taps} {MODE[mode]:20s} tile {tile} grid {grid}")
