import os, sys, time, torch
ROOT="/root/repo"
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "diffusion-handwriting-generation.pytorch_b200"))
from dhg_b200 import DiffusionWriter
from oracle import dhg_oracle as O
sd = O.init_state_dict(0)
torch.zeros(1).cuda(); torch.cuda.synchronize()
for dtype in ("bf16", "fp32"):
    t0 = time.perf_counter()
    w = DiffusionWriter(state_dict=sd, num_layers=2, channels=128, dtype=dtype)
    t1 = time.perf_counter()
    style = torch.randn(1, 14, 1280)
    out = w.sample(["Follow the White Rabbit"], style, seed=1).cpu()
    t2 = time.perf_counter()
    out = w.sample(["Follow the White Rabbit"], style, seed=1).cpu()
    t3 = time.perf_counter()
    print(f"{dtype}: construct+finalize {t1-t0:.2f} s, first sample (plan + tune + graph capture + chain) {t2-t1:.2f} s, second sample {1e3*(t3-t2):.1f} ms")
    w.close()
