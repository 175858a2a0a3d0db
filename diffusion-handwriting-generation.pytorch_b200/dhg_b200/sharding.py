"""Batch sharding of the sampling path across the GPUs of one box.

Every sample's chain is independent, so the path shards by
contiguous batch slices with NO collective inside the loop; the only exchange is
an optional gather of the finished [B,T,3] strokes at the end.
"""
import torch


def shard_bounds(total, rank, world):
    """Contiguous slice [lo, hi) of `total` samples owned by `rank`; sizes differ by at most 1."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError(f"bad rank/world {rank}/{world}")
    base, rem = divmod(int(total), world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def sample_sharded(sample_fn, text, style, x0, noise, rank, world, gather=True, group=None):
    """Run `sample_fn(text, style, x0=..., noise=...) -> [b,T,3]` on this rank's slice of a global
    batch and (optionally) gather the slices in rank order on every rank.

    `noise` is indexed [60, B, T, 2] by GLOBAL sample, so the result does not depend on `world`.
    """
    lo, hi = shard_bounds(text.shape[0], rank, world)
    if hi > lo:
        local = sample_fn(text[lo:hi], style[lo:hi], x0=x0[lo:hi], noise=noise[:, lo:hi])
    else:
        local = x0.new_zeros((0, x0.shape[1], 3))
    if not gather or world == 1:
        return local
    return gather_outputs(local, text.shape[0], rank, world, group=group)


def gather_outputs(local, total, rank, world, group=None):
    """All-gather the per-rank [b_r, T, 3] results into [total, T, 3] on every rank, in rank order.  One tensor
    collective (NCCL for CUDA tensors, gloo for CPU tensors): shards differ by at most one sample, so every rank
    contributes a buffer of the largest shard size and the padding row is dropped afterwards."""
    import torch.distributed as dist

    sizes = [hi - lo for lo, hi in (shard_bounds(total, r, world) for r in range(world))]
    width = max(sizes)
    send = local.new_zeros((width,) + tuple(local.shape[1:]))
    send[: local.shape[0]] = local
    recv = local.new_empty((world * width,) + tuple(local.shape[1:]))
    dist.all_gather_into_tensor(recv, send.contiguous(), group=group)
    recv = recv.view((world, width) + tuple(local.shape[1:]))
    return torch.cat([recv[r, : sizes[r]] for r in range(world)], dim=0)
