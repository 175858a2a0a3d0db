"""Elimination runs of the attention kernel (results are wrong by construction; timing only):
python tools/attn_elim.py H Tq Tk self masked [B] [D]   -- dbg bits: 1 no max pass, 2 no exp, 4 no P stores, 8 no O stores, 16 one PV MMA, 32 one S MMA"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "diffusion-handwriting-generation.pytorch_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
from test_gpu_attention import run_attention  # noqa: E402
from dhg_b200 import _abi  # noqa: E402

H, Tq, Tk, sa, mk = (int(x) for x in sys.argv[1:6])
B = int(sys.argv[6]) if len(sys.argv) > 6 else 1024
D = int(sys.argv[7]) if len(sys.argv) > 7 else 64
lib = _abi.lib()
out = []
for dbg in (0, 8, 4, 12, 2, 1, 3, 15, 48, 63):
    lib.dhg_set_option(None, b"attn_dbg", dbg)
    _, _, ms = run_attention(lib, B, H, D, Tq, Tk, bool(sa), bool(mk), 1, seed=1, repeats=5)
    out.append(f"dbg={dbg}: {ms * 1e3:.1f}")
lib.dhg_set_option(None, b"attn_dbg", 0)
print(f"H={H} Tq={Tq} Tk={Tk} self={sa} masked={mk} B={B} D={D} us:  " + "  ".join(out))
