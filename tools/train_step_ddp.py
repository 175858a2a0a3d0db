"""Data-parallel TRAINING step over NCCL (BASELINE configs[3]): every rank runs
DenoiserTrainer.train_step on its shard of a global batch (perturb, forward, loss, backward on this library's kernels),
the flat fp32 gradient (10,028,451 values) is summed over the ranks with ONE all-reduce and the 1/N is folded into the
clipped Adam kernel.  Checks, after two steps on a global batch of 8: (a) every rank holds identical parameters, (b) they
equal the parameters a single process gets from the whole batch (the mean of the shard losses is the global loss).
Then times the step at the reference's batch: 96 per rank (weak scaling) and 96 in total (96 / N per rank).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29541 tools/train_step_ddp.py
"""
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "diffusion-handwriting-generation.pytorch_b200"))
from dhg_b200.train import DenoiserTrainer  # noqa: E402
from oracle.dhg_oracle import init_state_dict  # noqa: E402  (seeded weights only)

rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", 0)))
if world > 1:
    dist.init_process_group("nccl")
sd = init_state_dict(0)


def batch(B, T, L, seed):
    g = torch.Generator().manual_seed(seed)
    text = torch.randint(2, 73, (B, L), generator=g)
    text[:, -1] = 1
    text[::2, L // 2:] = 0
    text[::2, L // 2 - 1] = 1
    return dict(strokes=torch.randn(B, T, 2, generator=g), pen=(torch.rand(B, T, generator=g) < 0.05).float(), text=text,
                style=torch.randn(B, 14, 1280, generator=g), keep=(torch.rand(B, 14, 1280, generator=g) >= 0.3).float() / 0.7,
                alphas=torch.rand(B, 1, generator=g) * 0.9 + 0.05, eps=torch.randn(B, T, 2, generator=g))


def run(tr, b, lo, hi, steps):
    c = {k: v[lo:hi].cuda().contiguous() for k, v in b.items()}
    out = None
    for _ in range(steps):
        out = tr.train_step(c["strokes"], c["pen"], c["text"], c["style"], c["alphas"], c["eps"], style_keep=c["keep"])
    return out


res = {"world": world}
# (a), (b): equivalence on a global batch of 8
G, T, L = 8, 32, 10
b = batch(G, T, L, 5)
per = G // world
tr = DenoiserTrainer(sd, per, T, L)
run(tr, b, rank * per, (rank + 1) * per, 2)
mine = tr.param.clone()
tr.close()
if world > 1:
    allp = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(allp, mine)
    res["max_diff_between_ranks"] = max((p - allp[0]).abs().max().item() for p in allp)
if rank == 0:
    # single process, whole batch: the exchange must be a no-op here, so the group is bypassed with group=None and world 1
    import dhg_b200.train as T_

    orig = T_.exchange_gradients
    T_.exchange_gradients = lambda flat_grad, group=None: 1
    try:
        one = DenoiserTrainer(sd, G, T, L)
        run(one, b, 0, G, 2)
        p0 = torch.cat([v.reshape(-1) for v in sd.values()]).cuda()
        res["rel_diff_vs_single_process"] = ((mine - one.param).norm() / (one.param - p0).norm()).item()   # relative to the size of the update
        one.close()
    finally:
        T_.exchange_gradients = orig
if world > 1:
    dist.barrier()
# timing at the reference's batch
for name, B in (("weak_96_per_rank", 96), ("global_96", 96 // world)):
    T, L = 480, 50
    tb = batch(B, T, L, 9 + rank)
    tr = DenoiserTrainer(sd, B, T, L)
    run(tr, tb, 0, B, 2)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    loss = run(tr, tb, 0, B, 5)
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / 5], device="cuda")
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    res[name] = {"ms_per_step": ms.item(), "samples_per_s": B * world / (ms.item() * 1e-3), "per_rank_batch": B, "loss_rank0": loss[0].item()}
    tr.close()
if rank == 0:
    print(json.dumps(res))
if world > 1:
    dist.destroy_process_group()
