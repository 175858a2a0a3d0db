"""Run one bf16 chain at another shape and report finiteness / time: python tools/shape_check.py B T L"""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "diffusion-handwriting-generation.pytorch_b200"))
from dhg_b200 import DiffusionWriter  # noqa: E402
from oracle.dhg_oracle import init_state_dict  # noqa: E402

B, T, L = (int(a) for a in sys.argv[1:4])
w = DiffusionWriter(state_dict=init_state_dict(0), num_layers=2, channels=128, dtype="bf16", chunk=B)
g = torch.Generator().manual_seed(3)
text = torch.randint(2, 73, (B, L), generator=g)
text[:, -1] = 1
text[::4, L // 2:] = 0
style = torch.randn(B, 14, 1280, generator=g).cuda()
x0 = torch.randn(B, T, 2, generator=g).cuda()
noise = torch.randn(60, B, T, 2, generator=g).cuda()
text = text.cuda()
out = w.sample(text, style, T=T, x0=x0, noise=noise)   # plans (tile tuning) + first chain
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(2):
    out = w.sample(text, style, T=T, x0=x0, noise=noise)
torch.cuda.synchronize()
dt = (time.perf_counter() - t0) / 2
print(f"B={B} T={T} L={L}: finite {bool(torch.isfinite(out).all())}, {dt * 1e3:.1f} ms per chain, {B / dt:.0f} lines/s, "
      f"{dt / 60 * 1e6:.0f} us per step, launches {w.last_launch_count}")
