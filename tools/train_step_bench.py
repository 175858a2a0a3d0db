"""BASELINE configs[3] on one GPU: this library's training step beside the reference's eager step (bench.train_step_leg).

    python tools/train_step_bench.py [B T L]            # JSON on stdout
    python tools/train_step_bench.py --one [B T L]      # ONE step of this library only (for an ncu launch list)
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "diffusion-handwriting-generation.pytorch_b200"))
import torch  # noqa: E402

import bench  # noqa: E402

args = [a for a in sys.argv[1:] if not a.startswith("--")]
B, T, L = (int(a) for a in args[:3]) if len(args) >= 3 else (96, 480, 50)
dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
if "--one" in sys.argv:
    from dhg_b200.train import DenoiserTrainer
    from oracle.dhg_oracle import init_state_dict

    g = torch.Generator().manual_seed(0)
    tr = DenoiserTrainer(init_state_dict(0), B, T, L, device=dev)
    text = torch.randint(2, 73, (B, L), generator=g)
    text[:, -1] = 1
    a = (torch.rand(B, 1, generator=g) * 0.9 + 0.05).to(dev)
    r = tr.train_step(torch.randn(B, T, 2, generator=g).to(dev), (torch.rand(B, T, generator=g) < 0.05).float().to(dev), text.to(dev),
                      torch.randn(B, 14, 1280, generator=g).to(dev), a, torch.randn(B, T, 2, generator=g).to(dev))
    torch.cuda.synchronize()
    print("loss", r[0].item(), "launches", tr.last_launch_count)
else:
    print(json.dumps(bench.train_step_leg(dev, B, T, L)))
