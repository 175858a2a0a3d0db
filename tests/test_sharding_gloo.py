"""N>1 host logic on CPU: two gloo ranks shard a global batch, run a stand-in sampler on their
slice, gather, and must reproduce the single-rank result exactly (no collective in the loop;
SURVEY.md 8e)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from dhg_b200.sharding import sample_sharded, shard_bounds


def fake_sampler(text, style, x0, noise):
    """Per-sample function of (text, style, x0, noise) only -- like the real chain."""
    x = x0 + noise.sum(0) * 0.01 + style.mean(dim=(1, 2))[:, None, None]
    pen = (text.float().mean(1) / 73.0)[:, None, None].expand(-1, x.shape[1], 1)
    return torch.cat((x, pen), dim=2)


def _inputs(B=7, T=16, L=5):
    g = torch.Generator().manual_seed(3)
    return (torch.randint(0, 73, (B, L), generator=g), torch.randn(B, 14, 1280, generator=g),
            torch.randn(B, T, 2, generator=g), torch.randn(60, B, T, 2, generator=g))


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    text, style, x0, noise = _inputs()
    out = sample_sharded(lambda t, s, x0, noise: fake_sampler(t, s, x0, noise), text, style, x0, noise, rank, world)
    lo, hi = shard_bounds(text.shape[0], rank, world)
    # max-over-ranks timing reduction used by bench.py
    t = torch.tensor([float(rank + 1)])
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    q.put((rank, out, (lo, hi), t.item()))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharded_sampling_matches_single_rank():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    text, style, x0, noise = _inputs()
    ref = fake_sampler(text, style, x0, noise)
    spans = sorted(r[2] for r in results)
    assert spans == [(0, 4), (4, 7)]
    for rank, out, _, tmax in results:
        assert torch.equal(out, ref)
        assert tmax == 2.0
