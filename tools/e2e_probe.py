"""Where does the time of DiffusionWriter.sample_host go?  python tools/e2e_probe.py [B] [calls]"""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "diffusion-handwriting-generation.pytorch_b200"))
from dhg_b200 import DiffusionWriter  # noqa: E402
from oracle.dhg_oracle import init_state_dict  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
calls = int(sys.argv[2]) if len(sys.argv) > 2 else 6
w = DiffusionWriter(state_dict=init_state_dict(0), num_layers=2, channels=128, dtype="bf16")
g = torch.Generator().manual_seed(0)
text = torch.randint(2, 73, (B, 24), generator=g)
text[:, -1] = 1
style = torch.randn(B, 14, 1280, generator=g).pin_memory()
x0 = torch.randn(B, 392, 2, generator=g).pin_memory()
noise = torch.randn(60, B, 392, 2, generator=g).pin_memory()
text = text.pin_memory()
for i in range(calls):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    out = w.sample_host(text, style, x0, noise)
    torch.cuda.synchronize()
    print(f"call {i}: {1e3 * (time.perf_counter() - t0):.2f} ms", flush=True)
t0 = time.perf_counter()
for i in range(5):
    o = torch.empty(B, 392, 3, dtype=torch.float32, pin_memory=True)
print(f"pinned empty x5: {1e3 * (time.perf_counter() - t0):.2f} ms")
