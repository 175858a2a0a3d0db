"""Time of the text side of a step, 1..n steps at once: python tools/text_side_bench.py [B]"""
import ctypes
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "diffusion-handwriting-generation.pytorch_b200"))
from dhg_b200 import DiffusionWriter  # noqa: E402
from oracle.dhg_oracle import init_state_dict  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
w = DiffusionWriter(state_dict=init_state_dict(0), num_layers=2, channels=128, dtype="bf16")
g = torch.Generator().manual_seed(0)
text = torch.randint(2, 73, (B, 24), generator=g)
text[:, -1] = 1
w.denoise(torch.randn(B, 392, 2, generator=g), text, torch.rand(B, 1, generator=g), torch.randn(B, 14, 1280, generator=g))
torch.cuda.synchronize()
for sets in (1, 2):
    ms = ctypes.c_float(0)
    rc = w._lib.dhg_debug_time_text(w._ctx, sets, 20, ctypes.byref(ms))
    assert rc == 0, w._lib.dhg_last_error().decode()
    print(f"text side, {sets} step(s) at once: {ms.value * 1e3:.1f} us per step")
