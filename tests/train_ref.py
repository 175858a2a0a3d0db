"""Shared by the CPU (host build) and GPU tests of the training step: the golden file of the reference's train_step
(tests/golden/make_golden_train.py), the inputs it was made from, and the comparison of a flat gradient with it."""
import importlib.util
import os

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))


def _maker():
    spec = importlib.util.spec_from_file_location("make_golden_train", os.path.join(HERE, "golden", "make_golden_train.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)   # imports nothing of the reference until its main() runs
    return m


def golden():
    m = _maker()
    z = np.load(os.path.join(HERE, "golden", "train_step_small.npz"))
    inp = m.draw_inputs()
    assert np.array_equal(inp["text"].numpy(), z["text"]) and np.array_equal(inp["alphas"].numpy(), z["alphas"]), "seeded inputs drifted"
    return z, inp, m.FULL


def loss_grads(eps, score_pred, pen_lifts, pen_pred, alphas):
    """d loss / d score_pred, d loss / d pen_lifts_pred of loss.py:27-37 (what dhg_train_loss writes), in torch."""
    B, T, _ = eps.shape
    n = float(B * T)
    y = pen_lifts.clamp(1e-7, 1 - 1e-7)
    g_s = -2.0 * (eps - score_pred) / n
    g_p = alphas.reshape(B, 1) / n * (pen_pred - y) / (pen_pred * (1 - pen_pred)).clamp_min(1e-12)
    return g_s, g_p


def check_gradients(grad_of, z, full, keys, tol=2e-4):
    """grad_of(key) -> tensor.  All 323 norms and the stored full gradients.  Tensors whose true gradient is zero (the
    key-projection biases: softmax is shift invariant) are compared on an absolute floor."""
    total = float(np.sqrt((z["grad_norms"] ** 2).sum()))
    worst = 0.0
    for i, k in enumerate(keys):
        want = float(z["grad_norms"][i])
        got = float(grad_of(k).double().norm())
        err = abs(got - want) / max(want, 1e-5 * total)
        worst = max(worst, err)
        assert err < tol, (k, got, want)
    for k in full:
        want = torch.from_numpy(z["grad/" + k])
        got = grad_of(k).cpu().reshape(want.shape)
        rel = float((got - want).norm() / want.norm())
        worst = max(worst, rel)
        assert rel < tol, (k, rel)
    return worst
