"""bf16 chain vs the fp32 oracle for the current DHG_OPTS: python tools/tail_accuracy.py [B]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "diffusion-handwriting-generation.pytorch_b200"))
from dhg_b200 import DiffusionWriter  # noqa: E402
from oracle import dhg_oracle as O  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
sd = O.init_state_dict(0)
g = torch.Generator().manual_seed(11)
text = torch.randint(2, 73, (B, 24), generator=g)
text[:, -1] = 1
text[::3, 17:] = 0
style = torch.randn(B, 14, 1280, generator=g)
x0 = torch.randn(B, 392, 2, generator=g)
noise = torch.randn(60, B, 392, 2, generator=g)
ref = O.reverse_chain(sd, text, style, x0, noise)
w = DiffusionWriter(state_dict=sd, num_layers=2, channels=128, dtype="bf16")
out = w.sample(text, style, T=392, x0=x0, noise=noise).cpu()
rel = ((out[..., :2] - ref[..., :2]).norm() / ref[..., :2].norm()).item()
dp = (out[..., 2] - ref[..., 2]).abs()
sure = (ref[..., 2] - 0.5).abs() > 0.02
agree = ((out[..., 2] > 0.5) == (ref[..., 2] > 0.5))[sure].float().mean().item()
print(f"DHG_OPTS={os.environ.get('DHG_OPTS', '')!r}: strokes rel-L2 {rel:.3e}, pen |dp| mean {dp.mean():.3e} max {dp.max():.3e}, agreement {agree:.4f}")
