"""Run one sampling chain WITHOUT the CUDA graph (every kernel is a separate launch) for ncu:
   python tools/profile_chain.py [B] [dtype]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "diffusion-handwriting-generation.pytorch_b200"))
from dhg_b200 import DiffusionWriter  # noqa: E402
from oracle.dhg_oracle import init_state_dict  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
dtype = sys.argv[2] if len(sys.argv) > 2 else "bf16"
w = DiffusionWriter(state_dict=init_state_dict(0), num_layers=2, channels=128, dtype=dtype, chunk=B, graph=0)
g = torch.Generator().manual_seed(0)
text = torch.randint(2, 73, (B, 24), generator=g)
text[:, -1] = 1
x0 = torch.randn(B, 392, 2, generator=g).cuda()
noise = torch.randn(60, B, 392, 2, generator=g).cuda()
style = torch.randn(B, 14, 1280, generator=g).cuda()
text = text.cuda()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
out = w.sample(text, style, x0=x0, noise=noise)
e1.record()
torch.cuda.synchronize()
print(f"B={B} {dtype}: chain (no graph) {e0.elapsed_time(e1):.1f} ms, launches {w.last_launch_count}, finite {bool(torch.isfinite(out).all())}")
