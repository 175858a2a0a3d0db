"""Forward + backward pass of the training step (csrc/train_step.cu) checked WITHOUT a GPU: the kernels are functors
over a flat index, and tests/hostsim_build.py compiles the same source with g++ (-DDHG_HOSTSIM) so that the identical
bodies and the identical tape run on the host.  Checked against (a) the golden file made by the unmodified reference
(model in train mode, loss_fn, backward: tests/golden/make_golden_train.py) and (b) torch autograd through the oracle
at a second shape.  The GPU twin of this file is tests/test_gpu_train_step.py."""
import ctypes
import os
import sys

import pytest
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import hostsim_build  # noqa: E402
import train_ref  # noqa: E402

from oracle import dhg_oracle as O  # noqa: E402


@pytest.fixture(scope="module")
def hs():
    return hostsim_build.lib()


def layout(h, num_layers, channels=128):
    out, i = {}, 0
    name = ctypes.create_string_buffer(256)
    off, num = ctypes.c_int64(), ctypes.c_int64()
    while h.dhg_trainer_param_info(num_layers, channels, i, name, 256, ctypes.byref(off), ctypes.byref(num)) == 0:
        out[name.value.decode()] = (off.value, num.value)
        i += 1
    return out


class HostTrainer:
    def __init__(self, h, sd, B, T, L, num_layers):
        self.h, self.lay = h, layout(h, num_layers)
        n = h.dhg_trainer_param_count(num_layers, 128)
        self.param, self.grad = torch.empty(n), torch.full((n,), 3.0)   # the backward must overwrite, not add to, old gradients
        for k, (o, m) in self.lay.items():
            self.param[o:o + m] = sd[k].reshape(-1)
        self.tr = ctypes.c_void_p()
        rc = h.dhg_trainer_create(0, num_layers, 128, B, T, L, self.param.data_ptr(), self.grad.data_ptr(), ctypes.byref(self.tr))
        assert rc == 0, h.dhg_trainer_last_error()
        self.B, self.T = B, T

    def forward(self, x, text, sigma, style, keep):
        score, pen = torch.empty(self.B, self.T, 2), torch.empty(self.B, self.T)
        rc = self.h.dhg_trainer_forward(self.tr, x.data_ptr(), text.data_ptr(), sigma.data_ptr(), style.data_ptr(),
                                        keep.data_ptr() if keep is not None else None, score.data_ptr(), pen.data_ptr(), None)
        assert rc == 0, self.h.dhg_trainer_last_error()
        return score, pen

    def backward(self, g_s, g_p):
        assert self.h.dhg_trainer_backward(self.tr, g_s.data_ptr(), g_p.data_ptr(), None) == 0
        return self.grad

    def grad_of(self, k):
        o, m = self.lay[k]
        return self.grad[o:o + m]

    def close(self):
        self.h.dhg_trainer_destroy(self.tr)


def test_flat_layout_is_the_checkpoint_key_order(hs):
    for nl in (2, 4):
        spec = O.state_dict_spec(nl, 128)
        lay = layout(hs, nl)
        assert list(lay) == [k for k, _ in spec]
        off = 0
        for k, shape in spec:
            n = 1
            for s in shape:
                n *= s
            assert lay[k] == (off, n)
            off += n
        assert hs.dhg_trainer_param_count(nl, 128) == off
    assert hs.dhg_trainer_param_count(2, 128) == 10_028_451


def test_forward_and_gradients_match_the_reference_training_step(hs, state_dict):
    z, inp, full = train_ref.golden()
    B, T, L = inp["strokes"].shape[1], inp["strokes"].shape[2], inp["text"].shape[2]
    tr = HostTrainer(hs, state_dict, B, T, L, 2)
    alphas, eps = inp["alphas"][0], inp["eps"][0]
    x_p = (torch.sqrt(alphas).unsqueeze(-1) * inp["strokes"][0] + torch.sqrt(1 - alphas).unsqueeze(-1) * eps).contiguous()
    score, pen = tr.forward(x_p, inp["text"][0].contiguous(), torch.sqrt(alphas).reshape(B).contiguous(), inp["style"][0].contiguous(),
                            inp["keep"][0].contiguous())
    want_s, want_p = torch.from_numpy(z["score_pred0"]), torch.from_numpy(z["pen_pred0"])
    assert (score - want_s).norm() / want_s.norm() < 1e-5
    assert (pen - want_p).abs().max() < 1e-5
    g_s, g_p = train_ref.loss_grads(eps, score, inp["pen_lifts"][0], pen, alphas)
    tr.backward(g_s.contiguous(), g_p.contiguous())
    worst = train_ref.check_gradients(tr.grad_of, z, full, list(tr.lay))
    print("worst gradient error vs the reference:", worst)
    tr.close()


def test_gradients_match_autograd_at_another_shape(hs):
    """One attention layer, an odd batch, T = 24 (3 rows at the deepest level), no dropout mask, random output gradients."""
    B, T, L, NL = 3, 24, 7, 1
    sd = O.init_state_dict(5, NL, 128)
    g = torch.Generator().manual_seed(11)
    x, style = torch.randn(B, T, 2, generator=g), torch.randn(B, 14, 1280, generator=g)
    text = torch.randint(2, 73, (B, L), generator=g)
    text[:, -1] = 1
    text[2, 2:] = 0
    sigma = torch.rand(B, generator=g) * 0.9 + 0.05
    tr = HostTrainer(hs, sd, B, T, L, NL)
    score, pen = tr.forward(x, text, sigma, style, None)
    sdr = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    eps_r, pen_r = O.denoiser_forward(sdr, x, text, sigma.reshape(B, 1), style, NL)
    assert (score - eps_r).norm() / eps_r.norm() < 1e-5 and (pen - pen_r).abs().max() < 1e-5
    g_s, g_p = torch.randn(B, T, 2, generator=g), torch.randn(B, T, generator=g)
    tr.backward(g_s, g_p)
    ((eps_r * g_s).sum() + (pen_r * g_p).sum()).backward()
    total = torch.sqrt(sum((v.grad.double() ** 2).sum() for v in sdr.values())).item()
    for k in tr.lay:
        want = sdr[k].grad.reshape(-1)
        err = (tr.grad_of(k) - want).norm().item() / max(want.norm().item(), 1e-5 * total)
        assert err < 2e-4, (k, err)
    tr.close()


def test_create_rejects_bad_plans(hs):
    buf = torch.zeros(hs.dhg_trainer_param_count(2, 128))
    tr = ctypes.c_void_p()
    assert hs.dhg_trainer_create(0, 2, 128, 2, 20, 5, buf.data_ptr(), buf.data_ptr(), ctypes.byref(tr)) != 0   # T not a multiple of 8
    assert b"multiple of 8" in hs.dhg_trainer_last_error()
    assert hs.dhg_trainer_create(0, 2, 128, 0, 16, 5, buf.data_ptr(), buf.data_ptr(), ctypes.byref(tr)) != 0
    assert hs.dhg_trainer_create(0, 2, 128, 2, 16, 5, None, buf.data_ptr(), ctypes.byref(tr)) != 0


@pytest.mark.parametrize("B,T,L,NL", [(1, 8, 1, 2), (2, 8, 3, 4)])
def test_smallest_shapes_and_four_attention_layers(hs, B, T, L, NL):
    """T = 8 leaves ONE row at the deepest level (a k3 convolution over a single row: two of its taps see only padding, and
    their weight gradients are exactly zero); L = 1 is a prompt of the end token alone; NL = 4 is DiffusionModel's own default."""
    sd = O.init_state_dict(9, NL, 128)
    g = torch.Generator().manual_seed(B * 100 + L)
    x, style = torch.randn(B, T, 2, generator=g), torch.randn(B, 14, 1280, generator=g)
    text = torch.randint(2, 73, (B, L), generator=g)
    text[:, -1] = 1
    sigma = torch.rand(B, generator=g) * 0.9 + 0.05
    tr = HostTrainer(hs, sd, B, T, L, NL)
    score, pen = tr.forward(x, text, sigma, style, None)
    sdr = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    eps_r, pen_r = O.denoiser_forward(sdr, x, text, sigma.reshape(B, 1), style, NL)
    assert (score - eps_r).norm() / eps_r.norm() < 1e-5 and (pen - pen_r).abs().max() < 1e-5
    g_s, g_p = torch.randn(B, T, 2, generator=g), torch.randn(B, T, generator=g)
    tr.backward(g_s, g_p)
    ((eps_r * g_s).sum() + (pen_r * g_p).sum()).backward()
    total = torch.sqrt(sum((v.grad.double() ** 2).sum() for v in sdr.values())).item()
    for k in tr.lay:
        want = sdr[k].grad.reshape(-1)
        # with a single key the softmax is constant: the query / key projections get an exactly zero gradient here and
        # rounding noise in autograd, hence the wider floor
        err = (tr.grad_of(k) - want).norm().item() / max(want.norm().item(), 1e-4 * total)
        assert err < 2e-4, (k, err)
    # calling the pair again gives the same gradient: the backward starts from zero, it does not accumulate across calls
    first = tr.grad.clone()
    tr.forward(x, text, sigma, style, None)
    tr.backward(g_s, g_p)
    assert (tr.grad - first).norm() / first.norm() < 1e-6
    tr.close()


def _shard_gradient(hs_lib, sd, inp, lo, hi):
    """Flat gradient of the loss of samples [lo, hi) (the loss is a mean over the shard, like a rank's own step)."""
    B, T, L = hi - lo, inp["x"].shape[1], inp["text"].shape[1]
    tr = HostTrainer(hs_lib, sd, B, T, L, 2)
    alphas, eps = inp["alphas"][lo:hi].contiguous(), inp["eps"][lo:hi].contiguous()
    score, pen = tr.forward(inp["x"][lo:hi].contiguous(), inp["text"][lo:hi].contiguous(), torch.sqrt(alphas).reshape(B).contiguous(),
                            inp["style"][lo:hi].contiguous(), None)
    g_s, g_p = train_ref.loss_grads(eps, score, inp["pen"][lo:hi], pen, alphas)
    out = tr.backward(g_s.contiguous(), g_p.contiguous()).clone()
    tr.close()
    return out


def _ddp_inputs():
    g = torch.Generator().manual_seed(77)
    G, T, L = 4, 16, 6
    text = torch.randint(2, 73, (G, L), generator=g)
    text[:, -1] = 1
    text[2, 3:] = 0
    return dict(x=torch.randn(G, T, 2, generator=g), text=text, style=torch.randn(G, 14, 1280, generator=g),
                pen=(torch.rand(G, T, generator=g) < 0.2).float(), alphas=torch.rand(G, 1, generator=g) * 0.9 + 0.05,
                eps=torch.randn(G, T, 2, generator=g))


def _ddp_rank(rank, world, port, out):
    import torch.distributed as dist

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), OMP_NUM_THREADS="4")
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from dhg_b200.train import exchange_gradients

    inp, sd = _ddp_inputs(), O.init_state_dict(0)
    per = inp["x"].shape[0] // world
    flat = _shard_gradient(hostsim_build.lib(), sd, inp, rank * per, (rank + 1) * per)
    n = exchange_gradients(flat)               # the step's one collective: SUM over the ranks
    out.put((rank, n, (flat / n).numpy()))     # the 1 / N the optimiser kernel folds in
    dist.destroy_process_group()


def test_data_parallel_gradient_equals_the_whole_batch_gradient(hs):
    """SURVEY 8e / BASELINE configs[3] on the host build, two gloo ranks: each rank's backward on its half of the batch,
    one SUM all-reduce of the flat gradient, divided by the world size = the gradient of the whole batch in one process."""
    import socket

    import torch.multiprocessing as mp

    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_ddp_rank, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted((q.get(timeout=300) for _ in procs), key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res[0][1] == 2 and (res[0][2] == res[1][2]).all()          # every rank ends with the same averaged gradient
    inp = _ddp_inputs()
    whole = _shard_gradient(hs, O.init_state_dict(0), inp, 0, inp["x"].shape[0])
    got = torch.from_numpy(res[0][2])
    assert (got - whole).norm() / whole.norm() < 1e-5
