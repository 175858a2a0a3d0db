"""The reference's training step (train.py:26-67; SURVEY.md 8a-18, 8e "optional train step", 8f-3) on the device, with the
reference's names and argument meaning.  The part around the model call:

    loss_fn(eps, score_pred, pen_lifts, pen_lifts_pred, alphas)        loss.py:5-39
    perturb(x, alphas, eps)                                            train.py:38-43
    InvSqrtSchedule(lr_mul, d_model, n_warmup_steps)                   scheduler.py:1-35 (the rate of step n)
    FlatAdam(params, lr_mul, d_model, n_warmup_steps, betas, weight_decay, clip_grad)
        .step_and_update_lr(flat_grad)                                 train.py:57-63: clip_grad_norm_ + Adam + schedule,
                                                                       and the data-parallel exchange in front of it

The kernels are CUDA (`csrc/train_update.cu`, C ABI `dhg_train_*`); torch tensors only carry the device memory, and
`torch.distributed` does the one collective of a data-parallel step (a SUM all-reduce of the flat fp32 gradient; the
division by the world size and the clip coefficient are folded into the optimiser kernel).

The forward + backward pass of the denoiser (`csrc/train_step.cu`, C ABI `dhg_trainer_*`) is `DenoiserTrainer`:

    DenoiserTrainer(state_dict, B, T, L).forward(x_perturbed, text, sigma, style)   model(...) in train.py:46-51
        .backward(grad_score, grad_pen)                                             loss.backward(), train.py:55
        .train_step(strokes, pen_lifts, text, style, alphas, eps)                   TrainingLoop.train_step, train.py:26-67
    get_alphas(batch_size, alpha_set)                                               utils/nn.py:42-61
    fit(trainer, batches, steps, exp_dir, ...)                                      the loop of TrainingLoop.train, train.py:95-137
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


def get_alphas(batch_size, alpha_set, generator=None):
    """utils/nn.py:42-61: one alpha per sample, uniform between two neighbouring entries of the alpha schedule.  Host
    side like the reference (its result is moved to the device by the caller); `generator` makes the draw repeatable."""
    idx = torch.randint(low=0, high=len(alpha_set) - 1, size=(batch_size, 1), dtype=torch.int64, generator=generator)
    lower, upper = alpha_set[idx], alpha_set[idx + 1]
    return torch.rand(lower.shape, generator=generator) * (upper - lower) + lower


def _trainer_check(rc):
    if rc != 0:
        raise DhgTrainError(_abi.lib().dhg_trainer_last_error().decode("utf-8", "replace"))


def param_layout(num_layers=2, channels=128):
    """Checkpoint key -> (offset, numel) of the flat parameter / gradient buffer (dhg_trainer_param_info): the 323 keys of
    model_final.pth in file order, tensors back to back with their checkpoint shapes."""
    from collections import OrderedDict

    lib = _abi.lib()
    name = ctypes.create_string_buffer(256)
    off, num = ctypes.c_int64(), ctypes.c_int64()
    out, i = OrderedDict(), 0
    while True:
        rc = lib.dhg_trainer_param_info(num_layers, channels, i, name, 256, ctypes.byref(off), ctypes.byref(num))
        if rc == -1:
            break
        _trainer_check(rc)
        out[name.value.decode()] = (off.value, num.value)
        i += 1
    return out


class DenoiserTrainer:
    """TrainingLoop.train_step (train.py:26-67) on the device: DiffusionModel.forward with every activation kept, the
    backward pass into ONE flat gradient buffer, and FlatAdam on the flat parameter buffer -- forward, loss, backward,
    gradient exchange and update are CUDA kernels of this library plus one NCCL all-reduce; nothing synchronises.

    `state_dict`: the checkpoint's tensors (any device); `B, T, L`: the plan's batch, stroke length (multiple of 8) and
    text length.  Optimiser arguments as in FlatAdam (config.yml:18-38)."""

    def __init__(self, state_dict, B, T, L, num_layers=None, channels=128, device="cuda", **optim):
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise DhgTrainError("DenoiserTrainer needs a CUDA device: there is no CPU path")
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        if num_layers is None:
            num_layers = len({k.split(".")[1] for k in state_dict if k.startswith("att_layers.")})
        self.num_layers, self.channels, self.B, self.T, self.L = num_layers, channels, B, T, L
        self.layout = param_layout(num_layers, channels)
        missing = [k for k in self.layout if k not in state_dict]
        extra = [k for k in state_dict if k not in self.layout]
        if missing or extra:   # strict, like load_state_dict(strict=True) (checkpoint.py:83-87)
            raise RuntimeError(f"state_dict does not match the model: missing {missing[:3]}, unexpected {extra[:3]}")
        for k, (_, n) in self.layout.items():
            if state_dict[k].numel() != n:
                raise RuntimeError(f"size mismatch for {k}: {tuple(state_dict[k].shape)}")
        self._shapes = {k: tuple(state_dict[k].shape) for k in self.layout}
        optim.setdefault("d_model", 2 * channels)
        self.optimizer = FlatAdam([state_dict[k].to(torch.float32) for k in self.layout], device=self.device, **optim)
        self.param = self.optimizer.param
        self.grad = torch.zeros_like(self.param)
        self._h = ctypes.c_void_p()
        _trainer_check(_abi.lib().dhg_trainer_create(self.device.index, num_layers, channels, B, T, L, _p(self.param), _p(self.grad),
                                                     ctypes.byref(self._h)))

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            _abi.lib().dhg_trainer_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:  # noqa: BLE001 -- interpreter shutdown
            pass

    @property
    def workspace_bytes(self):
        return _abi.lib().dhg_trainer_workspace_bytes(self._h)

    @property
    def last_launch_count(self):
        return _abi.lib().dhg_trainer_last_launches(self._h)

    def state_dict(self):
        """Views of the flat parameter buffer under the checkpoint's keys and shapes (what save_checkpoint writes)."""
        from collections import OrderedDict

        return OrderedDict((k, self.param[o:o + n].view(self._shapes[k])) for k, (o, n) in self.layout.items())

    def grad_dict(self):
        from collections import OrderedDict

        return OrderedDict((k, self.grad[o:o + n].view(self._shapes[k])) for k, (o, n) in self.layout.items())

    def forward(self, strokes, text, sigma, style_vector, style_keep=None):
        """model(x_perturbed, text, sigma, style) (model.py:121-182) -> (score_pred [B, T, 2], pen_lifts_pred [B, T], None).
        sigma: [B, 1] or [B, 1, 1]; text: integer [B, L]; style_keep: the Dropout(0.3) keep mask / 0.7 of text_style.py:92
        (None: no dropout)."""
        B, T, L = self.B, self.T, self.L
        strokes = _f32(strokes, (B, T, 2), "strokes")
        style_vector = _f32(style_vector, (B, 14, 1280), "style_vector")
        if style_keep is not None:
            style_keep = _f32(style_keep, (B, 14, 1280), "style_keep")
        sigma = _f32(sigma.reshape(B), (B,), "sigma")
        if not (isinstance(text, torch.Tensor) and text.is_cuda and tuple(text.shape) == (B, L)) or text.is_floating_point():
            raise ValueError(f"text: expected an integer CUDA tensor of shape {(B, L)}")
        text = text.to(torch.int64).contiguous()
        self._check_tokens(text)
        score = torch.empty(B, T, 2, dtype=torch.float32, device=self.device)
        pen = torch.empty(B, T, dtype=torch.float32, device=self.device)
        _trainer_check(_abi.lib().dhg_trainer_forward(self._h, _p(strokes), _p(text), _p(sigma), _p(style_vector), _p(style_keep), _p(score),
                                                      _p(pen), _stream()))
        return score, pen, None

    def _check_tokens(self, text):
        """Embedding(73, d) raises IndexError on an id outside [0, 73) (text_style.py:70).  The check reads one flag back from
        the device, so a tensor that was already checked and not written since (same object, same version) is not re-read."""
        key = (text.data_ptr(), text._version, tuple(text.shape))
        if getattr(self, "_tokens_ok", None) == key:
            return
        if bool(((text < 0) | (text >= 73)).any().item()):
            raise IndexError("text token id out of range [0, 73)")
        self._tokens_ok = key

    def backward(self, grad_score, grad_pen_pred):
        """loss.backward() (train.py:55) from d loss / d score_pred and d loss / d pen_lifts_pred: fills and returns the flat gradient."""
        B, T = self.B, self.T
        grad_score = _f32(grad_score, (B, T, 2), "grad_score")
        grad_pen_pred = _f32(grad_pen_pred.reshape(B, T), (B, T), "grad_pen_pred")
        _trainer_check(_abi.lib().dhg_trainer_backward(self._h, _p(grad_score), _p(grad_pen_pred), _stream()))
        return self.grad

    def train_step(self, strokes, pen_lifts, text, style_vectors, alphas, eps, style_keep=None, group=None):
        """TrainingLoop.train_step (train.py:26-67) with the two random draws (alphas: get_alphas, eps: randn_like) injected.
        Returns (loss, score_loss, pen_lifts_loss) as 0-dim CUDA tensors; the parameters are updated in place."""
        B = self.B
        alphas = alphas.reshape(B, 1)
        x_perturbed = perturb(strokes, alphas, eps)                                                 # train.py:40-43
        score_pred, pen_pred, _ = self.forward(x_perturbed, text, torch.sqrt(alphas), style_vectors, style_keep)   # :46-51
        loss, score_loss, pen_loss, g_s, g_p = loss_fn(eps, score_pred, pen_lifts, pen_pred, alphas, with_grads=True)   # :52-54
        self.backward(g_s, g_p)                                                                    # :55
        self.optimizer.step_and_update_lr(self.grad, group)                                        # :57-63
        return loss, score_loss, pen_loss


def fit(trainer, batches, steps, exp_dir, alpha_set=None, log_freq=5, save_freq=1000, logger=None, generator=None):
    """The loop of TrainingLoop.train (train.py:95-137) around `trainer.train_step`, with the reference's cadence and file
    names: a log line whenever (count + 1) % log_freq == 0 (averages since the last line), `checkpoint_<count + 1>.pth`
    whenever (count + 1) % save_freq == 0, `model_final.pth` after `steps` steps; on KeyboardInterrupt `checkpoint_last.pth`
    and `model_last.pth`.  `batches`: an iterable of dicts with "strokes" [B, T, 3], "text" [B, L], "style" [B, 14, 1280]
    (train.py:69-85 `process_batch`); it is restarted when exhausted.  The per-step draws (alphas: utils/nn.py:42-61, eps,
    the Dropout(0.3) keep mask of text_style.py:92) come from `generator` (None: torch's global one).
    Returns the list of logged (step, loss, score_loss, pen_lifts_loss)."""
    import os
    import time

    from .checkpoint import save_checkpoint, save_model_final
    from .diffusion import get_alpha_bar

    if alpha_set is None:
        alpha_set = torch.as_tensor(get_alpha_bar(), dtype=torch.float32)
    say = logger.info if logger is not None else (lambda msg: None)
    dev = getattr(trainer, "device", None)
    to = (lambda t: t.to(dev, non_blocking=True)) if dev is not None else (lambda t: t)
    history, acc, count, start = [], [], 0, time.time()
    it = iter(batches)
    try:
        while True:
            try:
                batch = next(it)
            except StopIteration:
                it = iter(batches)
                batch = next(it)
            count += 1
            strokes, text, style = batch["strokes"], batch["text"], batch["style"]
            x, pen_lifts = strokes[:, :, :2].contiguous(), strokes[:, :, 2].contiguous()            # train.py:76
            alphas = get_alphas(len(x), alpha_set, generator=generator)                              # train.py:36
            eps = torch.randn(x.shape, generator=generator)                                          # train.py:37
            keep = (torch.rand(style.shape, generator=generator) >= 0.3).float() / 0.7               # text_style.py:83,92
            losses = trainer.train_step(to(x), to(pen_lifts), to(text), to(style), to(alphas), to(eps), style_keep=to(keep))
            acc.append([float(v) for v in losses])                                                   # train.py:65-67 (.item())
            if (count + 1) % log_freq == 0:
                mean = [sum(c) / len(acc) for c in zip(*acc)]
                say(f"Step {count + 1} | Loss: {mean[0]:.3f} | Score: {mean[1]:.3f} | Pen: {mean[2]:.3f} | Time: {time.time() - start:.3f} sec")
                history.append((count + 1, *mean))
                acc = []
            if (count + 1) % save_freq == 0:
                say("Saving checkpoint...")
                save_checkpoint(trainer, os.path.join(exp_dir, f"checkpoint_{count + 1}.pth"))
            if count >= steps:
                say("Training finished, saving model weights.")
                save_model_final(trainer, os.path.join(exp_dir, "model_final.pth"))
                break
    except KeyboardInterrupt:
        say("Training interrupted by user.")
        save_checkpoint(trainer, os.path.join(exp_dir, "checkpoint_last.pth"))
        save_model_final(trainer, os.path.join(exp_dir, "model_last.pth"))
    return history
