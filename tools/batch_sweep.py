"""Device-resident chain time against the batch of one captured chain:  python tools/batch_sweep.py [dtype] [B ...]
(is a smaller chunk whose activations stay in the 126 MB L2 faster per prompt than B = 1024?)"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "diffusion-handwriting-generation.pytorch_b200"))
from dhg_b200 import DiffusionWriter  # noqa: E402
from oracle.dhg_oracle import init_state_dict  # noqa: E402

dtype = sys.argv[1] if len(sys.argv) > 1 else "bf16"
Bs = [int(x) for x in sys.argv[2:]] or [128, 256, 384, 512, 768, 1024]
sd = init_state_dict(0)
for B in Bs:
    w = DiffusionWriter(state_dict=sd, num_layers=2, channels=128, dtype=dtype, chunk=B)
    g = torch.Generator().manual_seed(0)
    text = torch.randint(2, 73, (B, 24), generator=g)
    text[:, -1] = 1
    text = text.cuda()
    x0 = torch.randn(B, 392, 2, generator=g).cuda()
    noise = torch.randn(60, B, 392, 2, generator=g).cuda()
    style = torch.randn(B, 14, 1280, generator=g).cuda()
    for _ in range(2):
        w.sample(text, style, x0=x0, noise=noise)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 3
    e0.record()
    for _ in range(n):
        w.sample(text, style, x0=x0, noise=noise)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    print(f"B={B:5d} {dtype}: {ms / 60 * 1e3:8.1f} us per step  {B / ms * 1e3:8.0f} lines/s  {ms / 60 * 1e3 / B:6.3f} us per step and prompt", flush=True)
    del w
    torch.cuda.empty_cache()
