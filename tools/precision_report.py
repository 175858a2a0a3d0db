"""Accuracy of every precision mode against the reference-generated goldens and the CPU oracle (run on a B200):
fp32 (tcgen05, split storage), fp32_simt (CUDA cores, fp32 storage), bf16.  Also times the chain per mode at B = 1024.

    python tools/precision_report.py [--time]
"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "diffusion-handwriting-generation.pytorch_b200")):
    sys.path.insert(0, p)
from dhg_b200 import DiffusionWriter  # noqa: E402
from oracle import dhg_oracle as O  # noqa: E402


def rel(a, b):
    return ((a.double() - b.double()).norm() / b.double().norm()).item()


def pen(out, ref, margin):
    sure = (ref - 0.5).abs() > margin
    return ((out > 0.5) == (ref > 0.5))[sure].float().mean().item()


def main():
    sd = O.init_state_dict(0)
    modes = {"fp32": dict(dtype="fp32"), "fp32_simt": dict(dtype="fp32", gemm=0), "bf16": dict(dtype="bf16")}
    gold = lambda n: np.load(os.path.join(ROOT, "tests", "golden", n + ".npz"))   # noqa: E731
    for name, kw in modes.items():
        w = DiffusionWriter(state_dict=sd, num_layers=2, channels=128, **kw)
        for case in ("fwd_small", "fwd_reftest"):
            g = gold(case)
            t = lambda k: torch.tensor(g[k])   # noqa: E731
            eps, p, _ = w.denoise(t("strokes"), t("text"), t("sigma"), t("style"))
            print(f"{name:10s} {case:12s} eps rel {rel(eps.cpu(), t('eps')):.3e}  pen max abs {(p.cpu() - t('pen')).abs().max().item():.3e}")
        g = gold("fwd_small")
        t = lambda k: torch.tensor(g[k])   # noqa: E731
        taps = {}
        O.denoiser_forward(sd, t("strokes"), t("text"), t("sigma"), t("style"), taps=taps)
        w.denoise(t("strokes"), t("text"), t("sigma"), t("style"))
        worst = max((rel(w.debug_read(n).reshape(taps[n].shape), taps[n]), n) for n in ("h1", "h2c", "h2", "h3c", "h3", "att_in", "att0", "att1", "d3", "d2", "d1"))
        print(f"{name:10s} worst tapped activation rel {worst[0]:.3e} ({worst[1]})")
        for case, key in (("chain_c1", "out_new"), ("chain_small", "out_new"), ("chain_small", "out_standard")):
            g = gold(case)
            t = lambda k: torch.tensor(g[k])   # noqa: E731
            out = w.sample(t("text"), t("style"), x0=t("x0"), noise=t("noise"), diffusion_mode=key.split("_", 1)[1]).cpu()
            ref = t(key)
            print(f"{name:10s} {case:12s} {key:12s} strokes rel {rel(out[..., :2], ref[..., :2]):.3e}  pen agree @0/0.01/0.02 "
                  f"{pen(out[..., 2], ref[..., 2], 0):.4f} {pen(out[..., 2], ref[..., 2], 0.01):.4f} {pen(out[..., 2], ref[..., 2], 0.02):.4f}  "
                  f"|dp| mean {(out[..., 2] - ref[..., 2]).abs().mean().item():.2e}")
        w.close()
    if "--time" in sys.argv:
        B, T, L = 1024, 392, 24
        gen = torch.Generator().manual_seed(1)
        text = torch.randint(2, 73, (B, L), generator=gen)
        text[:, -1] = 1
        style, x0, noise = torch.randn(B, 14, 1280, generator=gen), torch.randn(B, T, 2, generator=gen), torch.randn(60, B, T, 2, generator=gen)
        ins = [x.cuda() for x in (text, style, x0, noise)]
        for name, kw in modes.items():
            w = DiffusionWriter(state_dict=sd, num_layers=2, channels=128, **kw)
            for _ in range(2):
                w.sample(ins[0], ins[1], x0=ins[2], noise=ins[3])
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            w.sample(ins[0], ins[1], x0=ins[2], noise=ins[3])
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            print(f"{name:10s} B=1024 chain {dt * 1e3:.1f} ms  {B / dt:.0f} lines/s  {dt / 60 * 1e6:.0f} us/step  launches {w.last_launch_count}")
            w.close()


if __name__ == "__main__":
    main()
