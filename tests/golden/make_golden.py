"""Generate the golden vectors under tests/golden/ by running the UNMODIFIED
reference (imported from /root/reference) in the build container.

    python tests/golden/make_golden.py

The reference cannot travel to the GPU box, so its outputs are committed as
small .npz fixtures.  Weights are NOT stored (40 MB): they come from the
portable seeded initialiser `oracle.dhg_oracle.init_state_dict`, loaded into
the reference `DiffusionModel` with strict=True (which also pins the 323-key
checkpoint layout).  Everything the reference draws from the global RNG
(`torch.randn_like` inside the posterior updates) is injected.

The loop around the model re-states inference.py:81-96 because that module
itself needs `fire`, which is not installed; the model, the posterior-update
functions, the beta schedule and the tokenizer are the reference's own.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")

from diffusion_handwriting_generation.model import DiffusionModel  # noqa: E402
from diffusion_handwriting_generation.tokenizer import Tokenizer  # noqa: E402
from diffusion_handwriting_generation.utils import nn as ref_nn  # noqa: E402

from oracle.dhg_oracle import init_state_dict, stroke_length  # noqa: E402


def build_reference(seed):
    sd = init_state_dict(seed)
    model = DiffusionModel(num_layers=2, c1=128, c2=192, c3=256, drop_rate=0.0)
    model.load_state_dict(sd, strict=True)
    return model.eval()


class InjectedNoise:
    """Patch torch.randn_like so that the reference's posterior update consumes
    noise[i] at loop index i."""

    def __init__(self, noise):
        self.noise = noise
        self.i = None
        self._orig = torch.randn_like

    def __enter__(self):
        def fake(x, *a, **k):
            return self.noise[self.i].to(x.dtype)

        torch.randn_like = fake
        return self

    def __exit__(self, *exc):
        torch.randn_like = self._orig


@torch.no_grad()
def reference_chain(model, text, style, x0, noise, mode):
    beta_set = ref_nn.get_beta_set()
    alpha_set = torch.cumprod(1 - beta_set, dim=0)
    bs = text.shape[0]
    x = x0.clone()
    with InjectedNoise(noise) as inj:
        for i in range(len(beta_set) - 1, -1, -1):
            inj.i = i
            alpha = alpha_set[i] * torch.ones((bs, 1, 1))
            beta = beta_set[i] * torch.ones((bs, 1, 1))
            a_next = alpha_set[i - 1] if i > 1 else torch.tensor(1.0)
            model_out, pen_lifts, _ = model(x, text, torch.sqrt(alpha), style)
            if mode == "standard":
                x = ref_nn.standard_diffusion_step(x, model_out, beta, alpha, add_sigma=bool(i))
            else:
                x = ref_nn.new_diffusion_step(x, model_out, beta, alpha, a_next)
    return torch.cat((x, pen_lifts.unsqueeze(2)), dim=2)


def rand_tokens(g, B, L, pad_tail):
    """ids in [2,72], end token 1, optional zero padding of the tail."""
    t = torch.randint(2, 73, (B, L), generator=g)
    for b in range(B):
        n = L - 1 - (pad_tail[b] if pad_tail else 0)
        t[b, n] = 1
        t[b, n + 1:] = 0
    return t


def main():
    torch.set_num_threads(8)
    tok = Tokenizer()
    model = build_reference(0)

    # --- schedule + tokenizer -------------------------------------------------
    beta = ref_nn.get_beta_set()
    prompts = ["Follow the White Rabbit", "", "Hello, World! 42?", "tab\there ~ ünï", "a" * 49]
    enc = {f"tok_{i}": np.array(tok.encode(p), dtype=np.int64) for i, p in enumerate(prompts)}
    np.savez_compressed(
        os.path.join(HERE, "schedule_tokenizer.npz"),
        beta=beta.numpy(), alpha_bar=torch.cumprod(1 - beta, 0).numpy(),
        prompts=np.array(prompts), **enc,
    )

    # --- single forwards ------------------------------------------------------
    g = torch.Generator().manual_seed(1234)
    # (a) small ragged case with padded text and per-sample sigma [B,1]
    B, T, L = 3, 64, 12
    text = rand_tokens(g, B, L, [0, 5, 9])
    strokes = torch.randn(B, T, 2, generator=g)
    sigma = torch.rand(B, 1, generator=g) * 0.9 + 0.1
    style = torch.randn(B, 14, 1280, generator=g)
    with torch.no_grad():
        eps, pen, _ = model(strokes, text, sigma, style)
    np.savez_compressed(
        os.path.join(HERE, "fwd_small.npz"), seed=0, text=text.numpy(), strokes=strokes.numpy(),
        sigma=sigma.numpy(), style=style.numpy(), eps=eps.numpy(), pen=pen.numpy(),
    )
    # (b) the shape family of the reference's own test (tests/test_model.py:6-21):
    #     T=400, L=40, 0/1 tokens, style [B,1,1280]
    B = 2
    strokes = torch.rand(B, 400, 2, generator=g)
    text = (torch.rand(B, 40, generator=g) < 0.25).long()
    sigma = torch.rand(B, 1, generator=g)
    style = torch.rand(B, 1, 1280, generator=g)
    with torch.no_grad():
        eps, pen, _ = model(strokes, text, sigma, style)
    np.savez_compressed(
        os.path.join(HERE, "fwd_reftest.npz"), seed=0, text=text.numpy(), strokes=strokes.numpy(),
        sigma=sigma.numpy(), style=style.numpy(), eps=eps.numpy(), pen=pen.numpy(),
    )

    # --- full chains ----------------------------------------------------------
    # (c) C1: batch 1, 'Follow the White Rabbit', T=392 (BASELINE configs[0])
    ids = tok.encode("Follow the White Rabbit")
    T = stroke_length(len(ids))
    assert T == 392
    text = torch.tensor([ids])
    style = torch.randn(1, 14, 1280, generator=g)
    x0 = torch.randn(1, T, 2, generator=g)
    noise = torch.randn(60, 1, T, 2, generator=g)
    out = reference_chain(model, text, style, x0, noise, "new")
    np.savez_compressed(
        os.path.join(HERE, "chain_c1.npz"), seed=0, text=text.numpy(), style=style.numpy(),
        x0=x0.numpy(), noise=noise.numpy(), out_new=out.numpy(),
    )
    # (d) small ragged batch, both diffusion modes
    B, T, L = 3, 72, 9
    text = rand_tokens(g, B, L, [0, 2, 6])
    style = torch.randn(B, 14, 1280, generator=g)
    x0 = torch.randn(B, T, 2, generator=g)
    noise = torch.randn(60, B, T, 2, generator=g)
    out_new = reference_chain(model, text, style, x0, noise, "new")
    out_std = reference_chain(model, text, style, x0, noise, "standard")
    np.savez_compressed(
        os.path.join(HERE, "chain_small.npz"), seed=0, text=text.numpy(), style=style.numpy(),
        x0=x0.numpy(), noise=noise.numpy(), out_new=out_new.numpy(), out_standard=out_std.numpy(),
    )
    for f in sorted(os.listdir(HERE)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(HERE, f)))


if __name__ == "__main__":
    main()
