"""The oracle (oracle/dhg_oracle.py) against the golden vectors produced by the
reference itself (tests/golden/make_golden.py).  CPU only."""
import numpy as np
import torch

from oracle import dhg_oracle as O


def _rel(a, b):
    return ((a - b).norm() / b.norm()).item()


def test_state_dict_layout(state_dict):
    # 323 fp32 tensors, 10,028,451 parameters (strict-loaded into the reference
    # DiffusionModel when the golden vectors were generated)
    assert len(state_dict) == 323
    assert sum(v.numel() for v in state_dict.values()) == 10_028_451
    assert torch.equal(state_dict["enc1.affine1.gamma_emb.bias"], torch.ones(64))


def test_schedule_matches_reference(golden):
    g = golden("schedule_tokenizer")
    beta = O.beta_schedule()
    assert np.array_equal(beta.numpy(), g["beta"])
    assert np.array_equal(O.alpha_bar(beta).numpy(), g["alpha_bar"])
    assert len(beta) == O.NUM_DIFFUSION_STEPS == 60


def test_forward_matches_reference_small(state_dict, golden):
    g = golden("fwd_small")
    eps, pen = O.denoiser_forward(state_dict, torch.tensor(g["strokes"]), torch.tensor(g["text"]),
                                  torch.tensor(g["sigma"]), torch.tensor(g["style"]))
    # same ATen ops in the same order: fp32 results agree to rounding noise
    assert (eps - torch.tensor(g["eps"])).abs().max() < 1e-5
    assert (pen - torch.tensor(g["pen"])).abs().max() < 1e-6


def test_forward_matches_reference_test_shapes(state_dict, golden):
    # the shape family of the reference's own tests/test_model.py (T=400, L=40, style [B,1,1280])
    g = golden("fwd_reftest")
    eps, pen = O.denoiser_forward(state_dict, torch.tensor(g["strokes"]), torch.tensor(g["text"]),
                                  torch.tensor(g["sigma"]), torch.tensor(g["style"]))
    assert eps.shape == (2, 400, 2) and pen.shape == (2, 400)
    assert (eps - torch.tensor(g["eps"])).abs().max() < 1e-5
    assert (pen - torch.tensor(g["pen"])).abs().max() < 1e-6


def test_chain_c1_matches_reference(state_dict, golden):
    g = golden("chain_c1")
    out = O.reverse_chain(state_dict, torch.tensor(g["text"]), torch.tensor(g["style"]),
                          torch.tensor(g["x0"]), torch.tensor(g["noise"]))
    ref = torch.tensor(g["out_new"])
    assert out.shape == (1, 392, 3)
    assert _rel(out[..., :2], ref[..., :2]) < 1e-5
    assert ((out[..., 2] > 0.5) == (ref[..., 2] > 0.5)).all()


def test_chain_small_both_modes(state_dict, golden):
    g = golden("chain_small")
    for mode in ("new", "standard"):
        out = O.reverse_chain(state_dict, torch.tensor(g["text"]), torch.tensor(g["style"]),
                              torch.tensor(g["x0"]), torch.tensor(g["noise"]), mode)
        ref = torch.tensor(g["out_" + mode])
        assert _rel(out[..., :2], ref[..., :2]) < 1e-5, mode
        assert ((out[..., 2] > 0.5) == (ref[..., 2] > 0.5)).float().mean() == 1.0


def test_fp64_truth_oracle_agrees(state_dict, golden):
    # fp64 oracle (mask follows the activation dtype) vs the reference's fp32 output: the chain is
    # well conditioned (rel-L2 2e-7)
    g = golden("fwd_small")
    sd64 = {k: v.double() for k, v in state_dict.items()}
    eps, pen = O.denoiser_forward(sd64, torch.tensor(g["strokes"]).double(), torch.tensor(g["text"]),
                                  torch.tensor(g["sigma"]).double(), torch.tensor(g["style"]).double())
    assert _rel(eps.float(), torch.tensor(g["eps"])) < 1e-5


def test_last_two_steps_add_no_noise(state_dict):
    # inference.py:87: alpha_next = 1 for i in {0,1} -> the draw is multiplied by 0
    abar = O.alpha_bar(O.beta_schedule())
    x = torch.randn(1, 8, 2)
    z = torch.randn(1, 8, 2)
    for i in (0, 1):
        a = abar[i] * torch.ones(1, 1, 1)
        b = O.beta_schedule()[i] * torch.ones(1, 1, 1)
        y0 = O.posterior_new(x, x * 0.1, b, a, torch.tensor(1.0), z)
        y1 = O.posterior_new(x, x * 0.1, b, a, torch.tensor(1.0), z * 0)
        assert torch.equal(y0, y1)
