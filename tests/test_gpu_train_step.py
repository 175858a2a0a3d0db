"""The training step on the GPU (csrc/train_step.cu + train_update.cu through dhg_b200.train.DenoiserTrainer) against the
golden file made by the UNMODIFIED reference's train step (tests/golden/make_golden_train.py: model in train mode, loss_fn,
backward, clip_grad_norm_, Adam inside InvSqrtScheduledOptim, two steps) and against torch autograd through the oracle.
Tolerances: predictions 1e-5, gradients 2e-4 per tensor (fp32, different summation orders, atomics), parameters after
two updates 1e-5 of their norm (Adam's first steps move every weight by about lr, whatever the gradient's size)."""
import os
import sys

import numpy as np
import pytest
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import train_ref  # noqa: E402

from oracle import dhg_oracle as O  # noqa: E402

pytestmark = pytest.mark.gpu


def _dev(t):
    return t.cuda().contiguous()


def test_forward_backward_match_the_reference_training_step(state_dict):
    from dhg_b200.train import DenoiserTrainer, loss_fn, perturb

    z, inp, full = train_ref.golden()
    B, T, L = inp["strokes"].shape[1], inp["strokes"].shape[2], inp["text"].shape[2]
    tr = DenoiserTrainer(state_dict, B, T, L)
    assert tr.param.numel() == 10_028_451 and list(tr.layout) == list(state_dict)
    alphas, eps = _dev(inp["alphas"][0]), _dev(inp["eps"][0])
    x_p = perturb(_dev(inp["strokes"][0]), alphas, eps)
    tr.grad.fill_(3.0)   # the backward overwrites old gradients
    score, pen, third = tr.forward(x_p, _dev(inp["text"][0]), torch.sqrt(alphas), _dev(inp["style"][0]), _dev(inp["keep"][0]))
    assert third is None and pen.shape == (B, T)
    want_s, want_p = torch.from_numpy(z["score_pred0"]), torch.from_numpy(z["pen_pred0"])
    assert (score.cpu() - want_s).norm() / want_s.norm() < 1e-5
    assert (pen.cpu() - want_p).abs().max() < 1e-5
    loss, s_loss, p_loss, g_s, g_p = loss_fn(eps, score, _dev(inp["pen_lifts"][0]), pen, alphas, with_grads=True)
    assert np.allclose([loss.item(), s_loss.item(), p_loss.item()], z["losses0"], rtol=1e-5)
    tr.backward(g_s, g_p)
    gd = {k: v.cpu() for k, v in tr.grad_dict().items()}
    worst = train_ref.check_gradients(lambda k: gd[k], z, full, list(tr.layout))
    print("worst gradient error vs the reference:", worst, "launches", tr.last_launch_count)
    # every kernel variant against the default: the per-thread bodies the host build checks (0), the smallest tile (2),
    # 3 x TF32 tensor-core products (3: same fp32 contract), plain TF32 products (4: torch's allow_tf32 precision class)
    from dhg_b200 import _abi

    flat = torch.cat([gd[k].reshape(-1) for k in tr.layout])
    for mode, tol in ((0, 1e-5), (2, 1e-5), (3, 5e-5), (4, 5e-3)):
        _abi.lib().dhg_trainer_set_option(b"tiled_gemm", mode)
        try:
            score2, pen2, _ = tr.forward(x_p, _dev(inp["text"][0]), torch.sqrt(alphas), _dev(inp["style"][0]), _dev(inp["keep"][0]))
            g2 = tr.backward(g_s, g_p).cpu()
        finally:
            _abi.lib().dhg_trainer_set_option(b"tiled_gemm", 1)
        err_s, err_g = ((score2 - score).norm() / score.norm()).item(), ((g2 - flat).norm() / flat.norm()).item()
        print(f"tiled_gemm={mode}: score {err_s:.2e}, flat gradient {err_g:.2e} from the default")
        assert err_s < tol and err_g < tol, (mode, err_s, err_g)
    tr.close()


def test_two_training_steps_match_the_reference(state_dict):
    from dhg_b200.train import DenoiserTrainer

    z, inp, full = train_ref.golden()
    S, B, T = inp["strokes"].shape[:3]
    L = inp["text"].shape[2]
    tr = DenoiserTrainer(state_dict, B, T, L, lr_mul=1.0, n_warmup_steps=10000, betas=(0.9, 0.98), weight_decay=1e-5, clip_grad=100.0)
    for s in range(S):
        losses = tr.train_step(_dev(inp["strokes"][s]), _dev(inp["pen_lifts"][s]), _dev(inp["text"][s]), _dev(inp["style"][s]),
                               _dev(inp["alphas"][s]), _dev(inp["eps"][s]), style_keep=_dev(inp["keep"][s]))
        assert np.allclose([v.item() for v in losses], z[f"losses{s}"], rtol=2e-5), (s, [v.item() for v in losses], z[f"losses{s}"])
    sd = {k: v.cpu() for k, v in tr.state_dict().items()}
    for i, k in enumerate(tr.layout):
        assert abs(sd[k].double().norm().item() - z["param_norms"][i]) <= 1e-5 * max(z["param_norms"][i], 1e-3), k
        delta = (sd[k] - state_dict[k]).double().norm().item()
        assert abs(delta - z["param_delta_norms"][i]) <= 2e-2 * z["param_delta_norms"][i] + 1e-9, (k, delta, z["param_delta_norms"][i])
    for k in full:
        want = torch.from_numpy(z["param/" + k])
        assert (sd[k] - want).norm() / want.norm() < 1e-5, k
    tr.close()


def test_gradients_match_autograd_at_a_larger_shape():
    """B = 4, T = 64, L = 12, padded prompts: the per-sample split of the weight gradients and every tile boundary of the GEMM."""
    from dhg_b200.train import DenoiserTrainer

    B, T, L = 4, 64, 12
    sd = O.init_state_dict(3)
    g = torch.Generator().manual_seed(21)
    x, style = torch.randn(B, T, 2, generator=g), torch.randn(B, 14, 1280, generator=g)
    text = torch.randint(2, 73, (B, L), generator=g)
    text[:, -1] = 1
    text[1, 5:] = 0
    text[3, 9:] = 0
    sigma = torch.rand(B, 1, generator=g) * 0.9 + 0.05
    g_s, g_p = torch.randn(B, T, 2, generator=g), torch.randn(B, T, generator=g)
    tr = DenoiserTrainer(sd, B, T, L)
    score, pen, _ = tr.forward(_dev(x), _dev(text), _dev(sigma), _dev(style))
    tr.backward(_dev(g_s), _dev(g_p))
    sdr = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    eps_r, pen_r = O.denoiser_forward(sdr, x, text, sigma, style)
    assert (score.cpu() - eps_r).norm() / eps_r.norm() < 1e-5 and (pen.cpu() - pen_r).abs().max() < 1e-5
    ((eps_r * g_s).sum() + (pen_r * g_p).sum()).backward()
    total = torch.sqrt(sum((v.grad.double() ** 2).sum() for v in sdr.values())).item()
    gd = tr.grad_dict()
    for k in tr.layout:
        want = sdr[k].grad
        err = (gd[k].cpu() - want).norm().item() / max(want.norm().item(), 1e-5 * total)
        assert err < 2e-4, (k, err)
    tr.close()


def test_trainer_error_behaviour(state_dict):
    from dhg_b200.train import DenoiserTrainer, DhgTrainError

    with pytest.raises(DhgTrainError, match="multiple of 8"):
        DenoiserTrainer(state_dict, 2, 20, 5)
    bad = dict(state_dict)
    bad.pop("output_dense.bias")
    with pytest.raises(RuntimeError, match="missing"):
        DenoiserTrainer(bad, 2, 16, 5)
    tr = DenoiserTrainer(state_dict, 2, 16, 5)
    with pytest.raises(ValueError):
        tr.forward(torch.zeros(2, 24, 2, device="cuda"), torch.ones(2, 5, dtype=torch.int64, device="cuda"), torch.ones(2, 1, device="cuda"),
                   torch.zeros(2, 14, 1280, device="cuda"))
    with pytest.raises(ValueError):
        tr.forward(torch.zeros(2, 16, 2, device="cuda"), torch.ones(2, 5, device="cuda"), torch.ones(2, 1, device="cuda"),
                   torch.zeros(2, 14, 1280, device="cuda"))
    with pytest.raises(IndexError):
        tr.forward(torch.zeros(2, 16, 2, device="cuda"), torch.full((2, 5), 73, dtype=torch.int64, device="cuda"), torch.ones(2, 1, device="cuda"),
                   torch.zeros(2, 14, 1280, device="cuda"))
    tr.close()
