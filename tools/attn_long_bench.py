"""Long-key self-attention (Tk > 256): CUDA-core kernel vs the tcgen05 key-block kernel, BASELINE configs[4] shapes."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "diffusion-handwriting-generation.pytorch_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
from dhg_b200 import _abi
from test_gpu_attention import run_attention
lib = _abi.lib()
for name, B, H, Tq in (("C5 L1 600x600 h3", 128, 3, 600), ("C5 L2 300x300 h4", 128, 4, 300)):
    res = []
    for impl in (0, 1):
        got, ref, ms = run_attention(lib, B, H, 64, Tq, Tq, True, False, impl, seed=1, repeats=5)
        fl = 4.0 * B * H * Tq * Tq * 64
        res.append(f"impl{impl}: {ms * 1e3:8.1f} us {fl / ms / 1e9:6.1f} TF/s err {(got - ref).abs().max().item():.2e}")
    print(f"{name:22s} " + "  ".join(res))
