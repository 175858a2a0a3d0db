"""The update half of the reference's training step (SURVEY.md 8e "optional train step", 8f-3) with the reference's
names and argument meaning:

    loss_fn(eps, score_pred, pen_lifts, pen_lifts_pred, alphas)        loss.py:5-39
    perturb(x, alphas, eps)                                            train.py:38-43
    InvSqrtSchedule(lr_mul, d_model, n_warmup_steps)                   scheduler.py:1-35 (the rate of step n)
    FlatAdam(params, lr_mul, d_model, n_warmup_steps, betas, weight_decay, clip_grad)
        .step_and_update_lr(flat_grad)                                 train.py:57-63: clip_grad_norm_ + Adam + schedule,
                                                                       and the data-parallel exchange in front of it

The kernels are CUDA (`csrc/train_update.cu`, C ABI `dhg_train_*`); torch tensors only carry the device memory, and
`torch.distributed` does the one collective of a data-parallel step (a SUM all-reduce of the flat fp32 gradient; the
division by the world size and the clip coefficient are folded into the optimiser kernel).  The backward pass of the
denoiser is NOT built: `step_and_update_lr` takes the flat gradient as an argument.
"""
import ctypes

import torch

from . import _abi


class DhgTrainError(RuntimeError):
    pass


def _check(rc):
    if rc != 0:
        raise DhgTrainError(_abi.lib().dhg_train_last_error().decode("utf-8", "replace"))


def _p(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p(0)


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _f32(t, shape, name):
    if not (isinstance(t, torch.Tensor) and t.is_cuda and t.dtype == torch.float32 and t.is_contiguous()):
        raise ValueError(f"{name}: expected a contiguous float32 CUDA tensor")
    if tuple(t.shape) != tuple(shape):
        raise ValueError(f"{name}: expected shape {tuple(shape)}, got {tuple(t.shape)}")
    return t


def _scratch(device):
    return torch.empty(_abi.lib().dhg_train_scratch_doubles(), dtype=torch.float64, device=device)


def perturb(x, alphas, eps):
    """train.py:38-43: sqrt(alphas)[..., None] * x + sqrt(1 - alphas)[..., None] * eps.  x, eps [B, T, 2]; alphas [B, 1] or [B]."""
    B, T, _ = x.shape
    x = _f32(x, (B, T, 2), "x")
    eps = _f32(eps, (B, T, 2), "eps")
    alphas = _f32(alphas.reshape(B), (B,), "alphas")
    out = torch.empty_like(x)
    _check(_abi.lib().dhg_train_perturb(x.device.index or 0, _p(x), _p(alphas), _p(eps), _p(out), B, T, _stream()))
    return out


def loss_fn(eps, score_pred, pen_lifts, pen_lifts_pred, alphas, with_grads=False):
    """loss.py:5-39: (total, score_loss, pen_lifts_loss) as 0-dim CUDA tensors (no host synchronisation).
    pen_lifts_pred may be [B, T] or the model's [B, T, 1].  with_grads=True also returns d total / d score_pred and
    d total / d pen_lifts_pred: the first step of the backward pass, fused into the same read of the inputs."""
    B, T, _ = eps.shape
    eps = _f32(eps, (B, T, 2), "eps")
    score_pred = _f32(score_pred, (B, T, 2), "score_pred")
    pen_lifts = _f32(pen_lifts, (B, T), "pen_lifts")
    pshape = tuple(pen_lifts_pred.shape)
    pen_lifts_pred = _f32(pen_lifts_pred.reshape(B, T), (B, T), "pen_lifts_pred")
    alphas = _f32(alphas.reshape(B), (B,), "alphas")
    losses = torch.empty(3, dtype=torch.float32, device=eps.device)
    g_s = torch.empty_like(score_pred) if with_grads else None
    g_p = torch.empty_like(pen_lifts_pred) if with_grads else None
    _check(_abi.lib().dhg_train_loss(eps.device.index or 0, _p(eps), _p(score_pred), _p(pen_lifts), _p(pen_lifts_pred), _p(alphas), B, T,
                                     _p(losses), _p(g_s), _p(g_p), _p(_scratch(eps.device)), _stream()))
    out = (losses[0], losses[1], losses[2])
    return out + (g_s, g_p.reshape(pshape)) if with_grads else out


class InvSqrtSchedule:
    """scheduler.py:22-35: lr(n) = lr_mul * d_model^-0.5 * min(n^-0.5, n * n_warmup_steps^-1.5), n = 1, 2, ..."""

    def __init__(self, lr_mul, d_model, n_warmup_steps):
        self.lr_mul, self.d_model, self.n_warmup_steps = lr_mul, d_model, n_warmup_steps

    def lr(self, n_steps):
        if n_steps < 1:
            raise ValueError("the schedule starts at step 1 (scheduler.py:31 increments before the first update)")
        return self.lr_mul * (self.d_model ** -0.5) * min(n_steps ** (-0.5), n_steps * self.n_warmup_steps ** (-1.5))


def flatten_params(tensors, device):
    """One contiguous fp32 buffer holding the tensors back to back, and the (offset, shape) table to get them back."""
    table, off = [], 0
    for t in tensors:
        table.append((off, tuple(t.shape)))
        off += t.numel()
    flat = torch.empty(off, dtype=torch.float32, device=device)
    for (o, shp), t in zip(table, tensors):
        flat[o:o + t.numel()].copy_(t.reshape(-1))
    return flat, table


def exchange_gradients(flat_grad, group=None):
    """The one collective of a data-parallel step: SUM all-reduce of the flat gradient over the ranks, in
    place.  Returns the world size; the division by it happens inside the optimiser kernel.  Without an initialised
    process group this is a no-op that returns 1."""
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()):
        return 1
    dist.all_reduce(flat_grad, op=dist.ReduceOp.SUM, group=group)
    return dist.get_world_size(group)


class FlatAdam:
    """InvSqrtScheduledOptim(torch.optim.Adam(params, lr, betas, weight_decay), lr_mul, d_model, n_warmup_steps) plus
    dispatch_clip_grad(..., clip_grad, mode="norm") on ONE flat fp32 parameter buffer (config.yml:18-38: betas (0.9, 0.98),
    weight_decay 1e-5, clip_grad 100, warmup 10000, d_model = 2 * channels, lr_mul = 1).

    `params`: the model's parameter tensors in `model.parameters()` order (they are copied into the flat buffer).
    `step_and_update_lr(flat_grad)`: flat_grad is this rank's gradient in the same layout; with an initialised
    torch.distributed group it is summed over the ranks first (the mean over the world is what gets clipped and applied,
    i.e. the gradient of the global batch).  Nothing in the step synchronises with the host."""

    def __init__(self, params, lr_mul=1.0, d_model=256, n_warmup_steps=10000, betas=(0.9, 0.98), eps=1e-8, weight_decay=1e-5,
                 clip_grad=100.0, device="cuda"):
        self.schedule = InvSqrtSchedule(lr_mul, d_model, n_warmup_steps)
        self.betas, self.eps, self.weight_decay, self.clip_grad = betas, eps, weight_decay, clip_grad
        self.device = torch.device(device)
        if self.device.type == "cuda" and self.device.index is None:   # "cuda" = the current device of this process (one rank per GPU)
            self.device = torch.device("cuda", torch.cuda.current_device())
        self.param, self.table = flatten_params([p.detach() for p in params], self.device)
        self.exp_avg = torch.zeros_like(self.param)
        self.exp_avg_sq = torch.zeros_like(self.param)
        self.n_steps = 0
        self._world = 1
        self._sq = torch.zeros(1, dtype=torch.float64, device=self.device)
        self._scratch = _scratch(self.device)

    @property
    def lr(self):
        return self.schedule.lr(max(self.n_steps, 1))

    def tensors(self):
        """Views of the flat buffer with the original shapes (for a state_dict / the forward pass)."""
        return [self.param[o:o + int(torch.Size(s).numel())].view(s) for o, s in self.table]

    def grad_norm(self):
        """Total 2-norm of the last (exchanged, averaged) gradient as a 0-dim CUDA tensor: what clip_grad_norm_ returns."""
        return self._sq.sqrt().float().reshape(()) / self._world

    def step_and_update_lr(self, flat_grad, group=None):
        flat_grad = _f32(flat_grad, self.param.shape, "flat_grad")
        lib, dev, st = _abi.lib(), self.device.index or 0, _stream()
        self._world = exchange_gradients(flat_grad, group)
        self.n_steps += 1                              # scheduler.py:31
        lr = self.schedule.lr(self.n_steps)
        sq = None
        if self.clip_grad is not None:                 # train.py:57-61
            _check(lib.dhg_train_sqnorm(dev, _p(flat_grad), flat_grad.numel(), _p(self._sq), _p(self._scratch), st))
            sq = self._sq
        _check(lib.dhg_train_adam_step(dev, _p(self.param), _p(flat_grad), _p(self.exp_avg), _p(self.exp_avg_sq), self.param.numel(),
                                       self.n_steps, lr, self.betas[0], self.betas[1], self.eps, self.weight_decay, _p(sq),
                                       float(self.clip_grad or 0.0), self._world, st))
        return lr
