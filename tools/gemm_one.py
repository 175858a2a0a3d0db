"""Run one GEMM family a few times (for ncu).  python tools/gemm_one.py rows K N taps [film] [res_post] [ln]"""
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
if "film" in flags:
    kw.update(film=1, raw=False, act=True)
if "res_post" in flags:
    kw.update(res_post=True, raw=True, act=False)
if "ln" in flags:
    kw.update(ln=True)
c = gemm_ref.make_case(rows, K, N, taps, seed=1, **kw)
ms = gemm_ref.run(_abi.lib(), c, repeats=3)
print(f"rows={rows} K={K} N={N} taps={taps} {flags}: {ms*1e3:.1f} us")
