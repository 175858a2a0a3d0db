"""Time the attention shapes of one denoiser step at batch B, both kernels.  python tools/attn_bench.py [B]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "diffusion-handwriting-generation.pytorch_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
from test_gpu_attention import run_attention  # noqa: E402
from dhg_b200 import _abi  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
lib = _abi.lib()
CASES = [("self L1 196x196 h3", 3, 196, 196, True, False, 1), ("self L2 98x98 h4", 4, 98, 98, True, False, 1),
         ("self L3 49x49 h6", 6, 49, 49, True, False, 2), ("cross L1 196x24 h3", 3, 196, 24, False, True, 1),
         ("cross L2 98x24 h4", 4, 98, 24, False, True, 1), ("cross L3 49x24 h6", 6, 49, 24, False, True, 2),
         ("text-style 24x70 h8 d48", 8, 24, 70, False, False, 1)]
tot = [0.0, 0.0]
if os.environ.get("ATTN_MAX_SLOTS"):
    lib.dhg_set_option(None, b"attn_max_slots", int(os.environ["ATTN_MAX_SLOTS"]))
IMPLS = (1,) if os.environ.get("ATTN_TC_ONLY") else (0, 1)
for name, H, Tq, Tk, sa, mk, cnt in CASES:
    line = f"{name:22s}"
    for impl in IMPLS:
        D = 48 if "d48" in name else 64
        got, ref, ms = run_attention(lib, B, H, D, Tq, Tk, sa, mk, impl, seed=1, repeats=5)
        fl = 4.0 * B * H * Tq * Tk * D
        line += f"  impl{impl}: {ms*1e3:8.1f} us {fl/ms/1e9:7.1f} TF/s err {(got-ref).abs().max().item():.2e}"
        tot[impl] += cnt * ms * 1e3
    print(line, flush=True)
print(f"per step: simt {tot[0]:.0f} us, tcgen05 {tot[1]:.0f} us")
