"""Golden vectors of the reference's TRAINING step (train.py:26-67), made by running the UNMODIFIED reference classes
imported from /root/reference in the build container:

    python tests/golden/make_golden_train.py      ->  tests/golden/train_step_small.npz

DiffusionModel in train() mode (drop_rate 0.0 as in data/best_exp/config.yml:25; the Dropout(0.3) on the style vectors,
text_style.py:83,92, IS active), loss.loss_fn, utils.clip_grad.dispatch_clip_grad(100, "norm"), torch.optim.Adam(lr 3e-4,
betas (0.9, 0.98), weight_decay 1e-5) inside scheduler.InvSqrtScheduledOptim(lr_mul 1, d_model 256, warmup 10000)
(train.py:140-155, config.yml:18-38).  train.py itself cannot be imported (addict / ruamel missing), so the twelve
lines of train_step are re-stated around the reference's own classes.  The three random draws of a step are injected:
alphas (utils/nn.py:42-61), eps (randn_like) and the dropout keep mask (torch.nn.functional.dropout is patched for the
p = 0.3 call only).  Weights come from the portable seeded initialiser (not stored).  Stored: the small inputs (the large ones are
re-drawn from the seed: the tests import draw_inputs from this file), the
predictions and losses of both steps, the norm of every one of the 323 gradients of step 1, a few full gradients, and
norms + a few full tensors of the parameters after two updates.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle.dhg_oracle import alpha_bar, beta_schedule, init_state_dict  # noqa: E402

FULL = ["input_dense.weight", "sigma_ffn.1.weight", "enc1.conv1.weight", "enc3.mha.wq.weight", "enc3.affine1.gamma_emb.weight",
        "text_style_model.emb.weight", "att_layers.1.ffn.3.bias", "dec1.fc.weight", "pen_lifts_dense.0.weight"]
B, T, L, STEPS = 2, 16, 6, 2


class InjectedDropout:
    def __init__(self):
        self.keep = None
        self._orig = torch.nn.functional.dropout

    def __enter__(self):
        def fake(x, p=0.5, training=True, inplace=False):
            if not training or p == 0.0:
                return x
            assert abs(p - 0.3) < 1e-12 and x.shape == self.keep.shape
            return x * self.keep

        torch.nn.functional.dropout = fake
        return self

    def __exit__(self, *exc):
        torch.nn.functional.dropout = self._orig


def draw_inputs():
    g = torch.Generator().manual_seed(4242)
    out = dict(
        strokes=torch.randn(STEPS, B, T, 2, generator=g),
        pen_lifts=(torch.rand(STEPS, B, T, generator=g) < 0.2).float(),
        text=torch.randint(2, 73, (STEPS, B, L), generator=g),
        style=torch.randn(STEPS, B, 14, 1280, generator=g),
        keep=(torch.rand(STEPS, B, 14, 1280, generator=g) >= 0.3).float() / 0.7,
        eps=torch.randn(STEPS, B, T, 2, generator=g),
    )
    out["text"][:, :, -1] = 1
    out["text"][:, 1, 4:] = 0    # a padded prompt: exercises the mask of the cross-attentions
    out["text"][:, 1, 3] = 1
    abar = alpha_bar(beta_schedule())
    idx = torch.randint(0, len(abar) - 1, (STEPS, B, 1), generator=g)
    out["alphas"] = torch.rand(STEPS, B, 1, generator=g) * (abar[idx + 1] - abar[idx]) + abar[idx]
    return out


def main():
    sys.path.insert(0, "/root/reference")
    from diffusion_handwriting_generation.loss import loss_fn
    from diffusion_handwriting_generation.model import DiffusionModel
    from diffusion_handwriting_generation.scheduler import InvSqrtScheduledOptim
    from diffusion_handwriting_generation.utils.clip_grad import dispatch_clip_grad

    torch.manual_seed(0)
    sd = init_state_dict(0)
    model = DiffusionModel(num_layers=2, c1=128, c2=192, c3=256, drop_rate=0.0)
    model.load_state_dict(sd, strict=True)
    model.train()
    opt = InvSqrtScheduledOptim(torch.optim.Adam(model.parameters(), lr=3e-4, weight_decay=1e-5, betas=(0.9, 0.98)), 1.0, 256, 10000)
    inp = draw_inputs()
    names = [n for n, _ in model.named_parameters()]
    assert names == list(sd.keys()), "model.parameters() order == checkpoint key order"
    out = {k: inp[k].numpy() for k in ("text", "alphas")}   # the rest is re-drawn from the seed by the tests (draw_inputs)
    with InjectedDropout() as drop:
        for s in range(STEPS):
            x, pen, text, style = inp["strokes"][s], inp["pen_lifts"][s], inp["text"][s], inp["style"][s]
            alphas, eps = inp["alphas"][s], inp["eps"][s]
            drop.keep = inp["keep"][s]
            x_perturbed = torch.sqrt(alphas).unsqueeze(-1) * x + torch.sqrt(1 - alphas).unsqueeze(-1) * eps   # train.py:40-43
            opt.zero_grad()
            score_pred, pen_pred, _ = model(x_perturbed, text, torch.sqrt(alphas), style)                   # :46-51
            loss, score_loss, pen_loss = loss_fn(eps, score_pred, pen, pen_pred, alphas)                    # :52-54
            loss.backward()                                                                                  # :55
            if s == 0:
                grads = {n: p.grad.detach().clone() for n, p in model.named_parameters()}
                out["grad_norms"] = np.array([grads[n].norm().item() for n in names], dtype=np.float64)
                for n in FULL:
                    out["grad/" + n] = grads[n].numpy()
            dispatch_clip_grad(model.parameters(), value=100.0)                                              # :57-61
            opt.step_and_update_lr()                                                                         # :63
            out[f"score_pred{s}"] = score_pred.detach().numpy()
            out[f"pen_pred{s}"] = pen_pred.detach().numpy()
            out[f"losses{s}"] = np.array([loss.item(), score_loss.item(), pen_loss.item()], dtype=np.float64)
    params = dict(model.named_parameters())
    out["param_norms"] = np.array([params[n].detach().norm().item() for n in names], dtype=np.float64)
    out["param_delta_norms"] = np.array([(params[n].detach() - sd[n]).norm().item() for n in names], dtype=np.float64)
    for n in FULL:
        out["param/" + n] = params[n].detach().numpy()
    np.savez_compressed(os.path.join(HERE, "train_step_small.npz"), **out)
    print("losses", out["losses0"], out["losses1"], "grad norm", float(np.sqrt((out["grad_norms"] ** 2).sum())))


if __name__ == "__main__":
    main()
