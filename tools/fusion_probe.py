import os, sys, time, torch
ROOT="/root/repo"
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "diffusion-handwriting-generation.pytorch_b200"))
from dhg_b200 import DiffusionWriter
from oracle import dhg_oracle as O
sd = O.init_state_dict(0)
# parity small
g = torch.Generator().manual_seed(7)
B, T, L = 4, 392, 24
text = torch.randint(2, 73, (B, L), generator=g); text[:, -1] = 1
style = torch.randn(B, 14, 1280, generator=g); x0 = torch.randn(B, T, 2, generator=g); noise = torch.randn(60, B, T, 2, generator=g)
ref = O.reverse_chain(sd, text, style, x0, noise)
w = DiffusionWriter(state_dict=sd, num_layers=2, channels=128, dtype="bf16")
out = w.sample(text, style, x0=x0, noise=noise).cpu()
rel = ((out[..., :2]-ref[..., :2]).norm()/ref[..., :2].norm()).item()
sure = (ref[...,2]-0.5).abs() > 1e-2
print("opts", os.environ.get("DHG_OPTS"), "rel", rel, "pen", ((out[...,2]>0.5)==(ref[...,2]>0.5))[sure].float().mean().item(), "launches", w.last_launch_count)
w.close()
B=1024
text = torch.randint(2, 73, (B, L), generator=g); text[:, -1] = 1
style = torch.randn(B, 14, 1280, generator=g).cuda(); x0 = torch.randn(B, T, 2, generator=g).cuda(); noise = torch.randn(60, B, T, 2, generator=g).cuda(); text=text.cuda()
w = DiffusionWriter(state_dict=sd, num_layers=2, channels=128, dtype="bf16")
for _ in range(3): w.sample(text, style, x0=x0, noise=noise)
torch.cuda.synchronize(); t0=time.perf_counter()
for _ in range(5): w.sample(text, style, x0=x0, noise=noise)
torch.cuda.synchronize(); dt=(time.perf_counter()-t0)/5
print("opts", os.environ.get("DHG_OPTS"), f"B=1024 {dt*1e3:.1f} ms/chain {B/dt:.0f} lines/s {dt/60*1e6:.0f} us/step launches {w.last_launch_count}")
