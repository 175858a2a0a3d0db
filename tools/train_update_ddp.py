"""Data-parallel update step over NCCL (optional train step): every rank holds its own gradient of the
reference-sized flat buffer (10,028,451 fp32), `FlatAdam.step_and_update_lr` sums them (one all-reduce), folds 1/N and the
clip coefficient into the optimiser kernel, and every rank must end with the parameters a single process gets from the
averaged gradient with torch.optim.Adam + clip_grad_norm_.  Also times all-reduce + clip + Adam on the device (max over ranks).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29533 tools/train_update_ddp.py
"""
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "diffusion-handwriting-generation.pytorch_b200"))
from dhg_b200.train import FlatAdam  # noqa: E402

rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", 0)))
if world > 1:
    dist.init_process_group("nccl")
n = 10_028_451
g0 = torch.Generator().manual_seed(0)
p0 = (0.05 * torch.randn(n, generator=g0)).cuda()
grads = []
for r in range(world):   # every rank can rebuild every rank's gradient: the reference result needs their mean
    g = torch.Generator().manual_seed(100 + r)
    grads.append(0.05 * torch.randn(n, generator=g))
opt = FlatAdam([p0], clip_grad=100.0)
steps = 3
for s in range(steps):
    opt.step_and_update_lr((grads[rank] * (s + 1)).cuda())
# single-process reference on the averaged gradient
ref = torch.nn.Parameter(p0.clone())
topt = torch.optim.Adam([ref], lr=3e-4, betas=(0.9, 0.98), weight_decay=1e-5)
mean = (sum(grads) / world).cuda()
for s in range(steps):
    ref.grad = mean * (s + 1)
    torch.nn.utils.clip_grad_norm_([ref], 100.0)
    lr = (256 ** -0.5) * min((s + 1) ** -0.5, (s + 1) * 10000 ** -1.5)
    for grp in topt.param_groups:
        grp["lr"] = lr
    topt.step()
err = ((opt.param - ref.detach()).abs().max() / ref.detach().abs().max()).item()
same = torch.tensor([float(opt.param.double().sum().item())], device="cuda", dtype=torch.float64)
if world > 1:
    lst = [torch.zeros_like(same) for _ in range(world)]
    dist.all_gather(lst, same)
    identical = all(x.item() == lst[0].item() for x in lst)
else:
    identical = True
# timing: exchange + clip + Adam, device events, max over ranks
gbuf = grads[rank].cuda()
for _ in range(3):
    opt.step_and_update_lr(gbuf)
if world > 1:
    dist.barrier()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    opt.step_and_update_lr(gbuf)
e1.record()
torch.cuda.synchronize()
us = torch.tensor([e0.elapsed_time(e1) * 1000 / 20], device="cuda")
if world > 1:
    dist.all_reduce(us, op=dist.ReduceOp.MAX)
if rank == 0:
    print(json.dumps({"n_gpus": world, "parameters": n, "steps_checked": steps, "max_rel_err_vs_torch_on_mean_gradient": err,
                      "ranks_hold_identical_parameters": identical, "us_per_exchange_clip_adam_step": us.item(),
                      "collective": "one all-reduce (SUM) of the flat fp32 gradient per step, NCCL" if world > 1 else "none (1 rank)"}))
if world > 1:
    dist.destroy_process_group()
assert err < 5e-5 and identical
