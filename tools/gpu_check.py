"""GPU bring-up check: per-module activations of one forward and full chains vs the CPU oracle.
Usage: python tools/gpu_check.py [fp32|bf16] [gemm=0|1]"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "diffusion-handwriting-generation.pytorch_b200"))
from dhg_b200 import DiffusionWriter  # noqa: E402
from oracle import dhg_oracle as O  # noqa: E402

dtype = sys.argv[1] if len(sys.argv) > 1 else "fp32"
gemm = int(sys.argv[2]) if len(sys.argv) > 2 else None
sd = O.init_state_dict(0)
w = DiffusionWriter(state_dict=sd, num_layers=2, channels=128, dtype=dtype, gemm=gemm)
print("writer up", dtype, "gemm", gemm, flush=True)


def rel(a, b):
    return ((a - b).norm() / b.norm()).item()


g = np.load(os.path.join(ROOT, "tests/golden/fwd_small.npz"))
strokes, text, sigma, style = (torch.tensor(g[k]) for k in ("strokes", "text", "sigma", "style"))
taps = {}
eps_o, pen_o = O.denoiser_forward(sd, strokes, text, sigma, style, taps=taps)
eps, pen, _ = w.denoise(strokes, text, sigma, style)
torch.cuda.synchronize()
B = strokes.shape[0]
for name in ["text_act", "h1", "h2c", "h2", "h3c", "h3", "att_in", "att0", "att1", "d3", "d2", "d1"]:
    got = w.debug_read(name)
    ref = taps["text" if name == "text_act" else name]
    if name == "text_act":
        ref = torch.nn.functional.silu(ref)
    print(f"  tap {name:9s} rel {rel(got.reshape(ref.shape), ref):.3e}", flush=True)
print(f"fwd_small: eps rel {rel(eps.cpu(), eps_o):.3e} max|d| {(eps.cpu()-eps_o).abs().max():.3e}  pen max|d| {(pen.cpu()-pen_o).abs().max():.3e}")
print(f"  vs golden: eps max|d| {(eps.cpu()-torch.tensor(g['eps'])).abs().max():.3e}", flush=True)

g = np.load(os.path.join(ROOT, "tests/golden/chain_small.npz"))
text, style, x0, noise = (torch.tensor(g[k]) for k in ("text", "style", "x0", "noise"))
for mode in ("new", "standard"):
    t0 = time.time()
    out = w.sample(text, style, x0=x0, noise=noise, diffusion_mode=mode).cpu()
    ref = torch.tensor(g["out_" + mode])
    agree = ((out[..., 2] > 0.5) == (ref[..., 2] > 0.5)).float().mean().item()
    print(f"chain_small {mode}: rel {rel(out[..., :2], ref[..., :2]):.3e} pen agree {agree:.4f} "
          f"launches {w.last_launch_count} ({time.time()-t0:.2f}s)", flush=True)
g = np.load(os.path.join(ROOT, "tests/golden/chain_c1.npz"))
text, style, x0, noise = (torch.tensor(g[k]) for k in ("text", "style", "x0", "noise"))
out = w.sample(text, style, x0=x0, noise=noise).cpu()
ref = torch.tensor(g["out_new"])
agree = ((out[..., 2] > 0.5) == (ref[..., 2] > 0.5)).float().mean().item()
print(f"chain_c1: rel {rel(out[..., :2], ref[..., :2]):.3e} pen agree {agree:.4f}", flush=True)

# quick timing at B=64 and B=256
for Bt in (64, 256):
    gen = torch.Generator().manual_seed(1)
    text = torch.randint(2, 73, (Bt, 24), generator=gen); text[:, -1] = 1
    style = torch.randn(Bt, 14, 1280, generator=gen)
    x0 = torch.randn(Bt, 392, 2, generator=gen).cuda(); noise = torch.randn(60, Bt, 392, 2, generator=gen).cuda()
    text = text.cuda(); style = style.cuda()
    w.sample(text, style, x0=x0, noise=noise)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); w.sample(text, style, x0=x0, noise=noise); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print(f"B={Bt} T=392 chain {ms:.1f} ms -> {Bt/ms*1e3:.1f} lines/s, {ms/60*1e3:.0f} us/step, plan {w.plan_bytes/1e6:.0f} MB", flush=True)
