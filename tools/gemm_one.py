"""Run one GEMM family a few times (for ncu / DHG_TRACE).
python tools/gemm_one.py rows K N taps [period=393] [film] [res_post] [res_pre] [ln] [rowbias] [act] [both] [dual=K2]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "diffusion-handwriting-generation.pytorch_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import gemm_ref  # noqa: E402
from dhg_b200 import _abi  # noqa: E402

rows, K, N, taps = (int(x) for x in sys.argv[1:5])
flags = sys.argv[5:]
kw = dict(period=393, pad_first=1)
for f in flags:
    if f.startswith("period="):
        kw["period"] = int(f[7:])
    if f.startswith("dual="):
        kw["dual_K2"] = int(f[5:])
if "film" in flags:
    kw.update(film=1)
if "act" in flags:
    kw.update(raw=False, act=True)
if "both" in flags:
    kw.update(raw=True, act=True)
for k in ("res_post", "res_pre", "ln", "rowbias"):
    if k in flags:
        kw[k] = True
c = gemm_ref.make_case(rows, K, N, taps, seed=1, **kw)
ms = gemm_ref.run(_abi.lib(), c, repeats=3)
print(f"rows={rows} K={K} N={N} taps={taps} {flags}: {ms*1e3:.1f} us")
