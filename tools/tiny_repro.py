"""Repro helper: a sequence of denoise shapes through ONE writer (re-planning in between), against the oracle.
python tools/tiny_repro.py MODE "B,T,L" "B,T,L" ..."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "diffusion-handwriting-generation.pytorch_b200"))
from dhg_b200 import DiffusionWriter
from oracle import dhg_oracle as O
sd = O.init_state_dict(0)
mode = sys.argv[1]
w = DiffusionWriter(state_dict=sd, num_layers=2, channels=128, dtype=mode)
for spec in sys.argv[2:]:
    B, T, L = (int(x) for x in spec.split(","))
    g = torch.Generator().manual_seed(T)
    strokes = torch.randn(B, T, 2, generator=g); text = torch.randint(2, 73, (B, L), generator=g); text[:, -1] = 1
    sigma = torch.rand(B, 1, generator=g) * 0.9 + 0.05; style = torch.randn(B, 14, 1280, generator=g)
    eps, pen, _ = w.denoise(strokes, text, sigma, style)
    torch.cuda.synchronize()
    eo, po = O.denoiser_forward(sd, strokes, text, sigma, style)
    print(mode, spec, "rel", ((eps.cpu() - eo).norm() / eo.norm()).item(), flush=True)
