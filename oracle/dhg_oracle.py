"""CPU oracle for the reverse-diffusion sampling hot path.

TEST INFRASTRUCTURE ONLY.  Nothing in the product package may import this file;
only `tests/`, `__graft_entry__.smoke()` and the `cpu_baseline` / `--impl
reference` legs of `bench.py` do, and there only as the checker / the timed CPU
baseline.  The product path is the CUDA library and fails loudly without it.

This is a functional restatement (state_dict in, tensors out; no nn.Module
tree) of the reference algorithm, written from the behaviour of

    diffusion_handwriting_generation/model.py:121-182      DiffusionModel.forward
    diffusion_handwriting_generation/model.py:35-58        EncoderLayer.forward
    diffusion_handwriting_generation/cnn.py:52-87          ConvBlock.forward
    diffusion_handwriting_generation/attention.py:15-23    PosEmbeddings
    diffusion_handwriting_generation/attention.py:26-87    scaled_dp_attn / MultiHeadAttention
    diffusion_handwriting_generation/conditioning.py:16-19 AffineTransformLayer (FiLM)
    diffusion_handwriting_generation/text_style.py:91-104  TextStyleEncoder.forward
    diffusion_handwriting_generation/utils/nn.py:19-39     beta schedule
    diffusion_handwriting_generation/utils/nn.py:64-112    posterior updates
    diffusion_handwriting_generation/inference.py:77-96    the 60-step loop

Parity pin: the reference ships no golden vectors (its single test asserts
nothing), so this oracle is pinned against OUTPUTS OF THE REFERENCE ITSELF, run
in the build container by `tests/golden/make_golden.py`, and committed under
`tests/golden/*.npz`.  `tests/test_oracle_golden.py` checks the pin on CPU.

Works in float32 (the reference's arithmetic) and float64 (truth oracle: here the
padding mask follows the activations' dtype, which avoids the fp64-q/fp32-mask
SDPA trap of running the reference module under `.double()`).
"""
from __future__ import annotations

import math
from collections import OrderedDict

import torch
import torch.nn.functional as F

NUM_DIFFUSION_STEPS = 60  # utils/nn.py:21 -- explin(1e-5, 0.4, 60)
SIGMA_DIM = 32            # conditioning.py:9 -- hard-coded Linear(32, hidden)
STYLE_SPLIT = 5           # text_style.py:92 -- reshape_up(style, 5)
VOCAB = 73                # text_style.py:70


# --------------------------------------------------------------------------
# state_dict layout (checkpoint.py:256-297 + train.py:131 => raw OrderedDict)
# --------------------------------------------------------------------------
def _affine_keys(prefix, hidden):
    return [
        (f"{prefix}.gamma_emb.weight", (hidden, SIGMA_DIM)),
        (f"{prefix}.gamma_emb.bias", (hidden,)),
        (f"{prefix}.beta_emb.weight", (hidden, SIGMA_DIM)),
        (f"{prefix}.beta_emb.bias", (hidden,)),
    ]


def _linear_keys(prefix, d_in, d_out):
    return [(f"{prefix}.weight", (d_out, d_in)), (f"{prefix}.bias", (d_out,))]


def _conv_keys(prefix, d_in, d_out):
    return [(f"{prefix}.weight", (d_out, d_in, 3)), (f"{prefix}.bias", (d_out,))]


def _mha_keys(prefix, d):
    out = []
    for n in ("wq", "wk", "wv", "dense"):
        out += _linear_keys(f"{prefix}.{n}", d, d)
    return out


def _conv_block_keys(prefix, d_in, d_out):
    out = []
    out += _affine_keys(f"{prefix}.affine1", d_out // 2)
    out += _affine_keys(f"{prefix}.affine2", d_out)
    out += _affine_keys(f"{prefix}.affine3", d_out)
    out += _conv_keys(f"{prefix}.conv_skip", d_in, d_out)
    out += _conv_keys(f"{prefix}.conv1", d_in, d_out // 2)
    out += _conv_keys(f"{prefix}.conv2", d_out // 2, d_out)
    out += _linear_keys(f"{prefix}.fc", d_out, d_out)
    return out


def _encoder_layer_keys(prefix, d_in, d_out):
    out = []
    out += _linear_keys(f"{prefix}.text_dense", d_in, d_out)
    out += _linear_keys(f"{prefix}.ffn.1", d_out, 2 * d_out)
    out += _linear_keys(f"{prefix}.ffn.3", 2 * d_out, d_out)
    out += _mha_keys(f"{prefix}.mha", d_out)
    out += _mha_keys(f"{prefix}.mha2", d_out)
    for i in range(4):
        out += _affine_keys(f"{prefix}.affine{i}", d_out)
    return out


def state_dict_spec(num_layers=2, channels=128):
    """Ordered (key, shape) list of `model_final.pth`."""
    c1, c2, c3 = channels, channels * 3 // 2, channels * 2
    d = 2 * c2
    spec = []
    spec += _linear_keys("input_dense", 2, c1)
    spec += _linear_keys("sigma_ffn.1", 1, 2048)
    spec += _linear_keys("sigma_ffn.3", 2048, c1 // 4)
    spec += _conv_block_keys("enc1", c1, c1)
    spec += _conv_block_keys("enc2", c1, c2)
    spec += _encoder_layer_keys("enc3", d, c2)
    spec += _conv_block_keys("enc4", c2, c3)
    spec += _encoder_layer_keys("enc5", d, c3)
    spec += _conv_keys("skip_conv1", c1, c2)
    spec += _conv_keys("skip_conv2", c2, c3)
    spec += _conv_keys("skip_conv3", c3, d)
    spec += [("text_style_model.emb.weight", (VOCAB, d))]
    spec += _linear_keys("text_style_model.style_ffn.1", 256, 4 * c2)
    spec += _linear_keys("text_style_model.style_ffn.3", 4 * c2, d)
    spec += _linear_keys("text_style_model.text_ffn.1", d, 2 * d)
    spec += _linear_keys("text_style_model.text_ffn.3", 2 * d, d)
    spec += _mha_keys("text_style_model.mha", d)
    for i in range(1, 5):
        spec += _affine_keys(f"text_style_model.affine{i}", d)
    spec += _linear_keys("att_dense", 2 * c1, d)
    for i in range(num_layers):
        spec += _encoder_layer_keys(f"att_layers.{i}", d, d)
    spec += _conv_block_keys("dec3", d, c3)
    spec += _conv_block_keys("dec2", c3, c2)
    spec += _conv_block_keys("dec1", c2, c1)
    spec += _linear_keys("output_dense", c1, 2)
    spec += _linear_keys("pen_lifts_dense.0", c1, 1)
    return spec


def init_state_dict(seed=0, num_layers=2, channels=128, dtype=torch.float32):
    """Portable seeded random init with the reference's init *distributions*
    (Linear/Conv: U(+-1/sqrt(fan_in)); Embedding: N(0,1); gamma bias = 1,
    conditioning.py:13).  Draw order = spec order, one CPU generator, so the
    GPU box regenerates bit-identical weights without the reference present."""
    g = torch.Generator(device="cpu")
    g.manual_seed(int(seed))
    sd = OrderedDict()
    fan = {}
    for key, shape in state_dict_spec(num_layers, channels):
        if key.endswith("emb.weight") and len(shape) == 2 and shape[0] == VOCAB:
            t = torch.randn(shape, generator=g, dtype=torch.float32)
        elif key.endswith(".weight"):
            fan_in = 1
            for s in shape[1:]:
                fan_in *= s
            fan[key[: -len(".weight")]] = fan_in
            bound = 1.0 / math.sqrt(fan_in)
            t = (torch.rand(shape, generator=g, dtype=torch.float32) * 2 - 1) * bound
        else:  # bias
            base = key[: -len(".bias")]
            bound = 1.0 / math.sqrt(fan[base])
            t = (torch.rand(shape, generator=g, dtype=torch.float32) * 2 - 1) * bound
            if base.endswith("gamma_emb"):
                t = torch.ones(shape, dtype=torch.float32)
        sd[key] = t.to(dtype)
    return sd


# --------------------------------------------------------------------------
# building blocks
# --------------------------------------------------------------------------
def _lin(sd, p, x):
    return F.linear(x, sd[p + ".weight"], sd[p + ".bias"])


def _conv3(sd, p, x_btc):
    """k=3, dilation 1, zero 'same' padding (cnn.py:32-47 -- dils[1] is never
    read).  Takes and returns channels-last [B,T,C]."""
    y = F.conv1d(x_btc.transpose(1, 2), sd[p + ".weight"], sd[p + ".bias"], padding=1)
    return y.transpose(1, 2)


def _film(sd, p, x, sig):
    """conditioning.py:16-19.  sig: [B,1,32]."""
    g = _lin(sd, p + ".gamma_emb", sig).reshape(sig.shape[0], 1, -1)
    b = _lin(sd, p + ".beta_emb", sig).reshape(sig.shape[0], 1, -1)
    return x * g + b


def _ln(x):
    """LayerNorm(eps=1e-6, elementwise_affine=False) (model.py:25)."""
    return F.layer_norm(x, (x.shape[-1],), eps=1e-6)


def _ffn(sd, p, x):
    """utils/nn.py:145-175, act_before=True: SiLU -> Linear -> SiLU -> Linear."""
    return _lin(sd, p + ".3", F.silu(_lin(sd, p + ".1", F.silu(x))))


def pos_table(length, dim, pos_factor, dtype=torch.float32):
    """attention.py:15-23: halves concatenated (sin | cos), not interleaved."""
    half = dim // 2
    step = math.log(10000) / (half - 1)
    freq = torch.exp(torch.arange(half) * -step)          # fp32 like the reference
    ang = torch.arange(length)[:, None] * freq[None, :] * pos_factor
    return torch.cat((ang.sin(), ang.cos()), dim=-1)[None].to(dtype)


def _mha(sd, p, q, k, v, heads, mask=None):
    """attention.py:49-87.  mask: [B,1,1,Lk] with 1 at padded keys; applied as
    an additive -1e9 (attention.py:43)."""
    B, d = q.shape[0], q.shape[-1]
    depth = d // heads

    def split(t):
        return t.reshape(B, -1, heads, depth).transpose(1, 2)

    qh, kh, vh = split(_lin(sd, p + ".wq", q)), split(_lin(sd, p + ".wk", k)), split(_lin(sd, p + ".wv", v))
    am = None if mask is None else mask.to(qh.dtype) * -1e9
    o = F.scaled_dot_product_attention(qh, kh, vh, attn_mask=am)
    o = o.transpose(1, 2).reshape(B, -1, d)
    return _lin(sd, p + ".dense", o)


def conv_block(sd, p, x, sig):
    """cnn.py:52-87 in channels-last form.  x: [B,T,Cin] -> [B,T,Cout]."""
    skip = _conv3(sd, p + ".conv_skip", x)
    y = _film(sd, p + ".affine1", _conv3(sd, p + ".conv1", F.silu(x)), sig)
    y = _film(sd, p + ".affine2", _conv3(sd, p + ".conv2", F.silu(y)), sig)
    y = _film(sd, p + ".affine3", _lin(sd, p + ".fc", F.silu(y)), sig)
    return y + skip


def encoder_layer(sd, p, x, text, sig, mask, heads, pos_factor):
    """model.py:35-58.  x: [B,T',d], text: [B,L,384]."""
    d = sd[p + ".text_dense.weight"].shape[0]
    t = _film(sd, p + ".affine0", _ln(_lin(sd, p + ".text_dense", F.silu(text))), sig)
    t_pe = t + pos_table(t.shape[1], d, 1.0, t.dtype)
    x_pos = pos_table(x.shape[1], d, pos_factor, x.dtype)
    x_pe = x + x_pos
    x2 = _mha(sd, p + ".mha", x_pe, t_pe, t, heads, mask)        # v carries no PE
    x2 = _film(sd, p + ".affine1", _ln(x2), sig) + x             # LN has no residual inside
    x2_pe = x2 + x_pos
    x3 = _mha(sd, p + ".mha2", x2_pe, x2_pe, x2, heads)          # v carries no PE
    x3 = _film(sd, p + ".affine2", _ln(x2 + x3), sig)
    x4 = _ffn(sd, p + ".ffn", x3) + x3
    return _film(sd, p + ".affine3", _ln(x4), sig)


def text_style_encoder(sd, text_ids, style, sig):
    """text_style.py:91-104 (eval mode: Dropout(0.3) inactive)."""
    p = "text_style_model"
    B, n, c = style.shape
    s = style.reshape(B, n * STYLE_SPLIT, c // STYLE_SPLIT)      # reshape_up, utils/nn.py:115-127
    s = _film(sd, p + ".affine1", _ln(_ffn(sd, p + ".style_ffn", s)), sig)
    t = F.embedding(text_ids.long(), sd[p + ".emb.weight"])
    t = _film(sd, p + ".affine2", _ln(t), sig)
    m = _mha(sd, p + ".mha", t, s, s, 8)                          # no mask here
    t = _film(sd, p + ".affine3", _ln(t + m), sig)
    return _film(sd, p + ".affine4", _ln(_ffn(sd, p + ".text_ffn", t)), sig)  # no residual


def _pool(x):   # AvgPool1d(2) over T, channels-last
    B, T, C = x.shape
    return x.reshape(B, T // 2, 2, C).mean(dim=2)


def _up(x):     # nearest x2 over T, channels-last
    return x.repeat_interleave(2, dim=1)


def sigma_embed(sd, sigma):
    """sigma_ffn (model.py:83,134).  sigma [B,1,1] or [B,1] -> [B,1,32]."""
    s = sigma.reshape(sigma.shape[0], 1, 1)
    return _ffn(sd, "sigma_ffn", s)


def denoiser_forward(sd, strokes, text_ids, sigma, style, num_layers=None, taps=None):
    """DiffusionModel.forward (model.py:121-182), channels-last throughout.
    Returns (eps [B,T,2], pen [B,T]).  `taps`, if a dict, receives named
    intermediate activations for kernel-level debugging."""
    dt = strokes.dtype
    if num_layers is None:
        num_layers = len({k.split(".")[1] for k in sd if k.startswith("att_layers.")})
    sig = sigma_embed(sd, sigma.to(dt))
    mask = (text_ids == 0).to(dt)[:, None, None, :]              # utils/nn.py:189-191
    text = text_style_encoder(sd, text_ids, style.to(dt), sig)

    def tap(name, v):
        if taps is not None:
            taps[name] = v.detach().clone()
        return v

    tap("sig", sig)
    tap("text", text)
    x = _lin(sd, "input_dense", strokes)
    h1 = tap("h1", conv_block(sd, "enc1", x, sig))
    h2 = tap("h2c", conv_block(sd, "enc2", _pool(h1), sig))
    h2 = tap("h2", encoder_layer(sd, "enc3", h2, text, sig, mask, 3, 4.0))
    h3 = tap("h3c", conv_block(sd, "enc4", _pool(h2), sig))
    h3 = tap("h3", encoder_layer(sd, "enc5", h3, text, sig, mask, 4, 2.0))
    x = tap("att_in", _lin(sd, "att_dense", _pool(h3)))
    for i in range(num_layers):
        x = tap(f"att{i}", encoder_layer(sd, f"att_layers.{i}", x, text, sig, mask, 6, 1.0))
    x = tap("d3", conv_block(sd, "dec3", _up(x) + _conv3(sd, "skip_conv3", h3), sig))
    x = tap("d2", conv_block(sd, "dec2", _up(x) + _conv3(sd, "skip_conv2", h2), sig))
    x = tap("d1", conv_block(sd, "dec1", _up(x) + _conv3(sd, "skip_conv1", h1), sig))
    eps = _lin(sd, "output_dense", x)
    pen = torch.sigmoid(_lin(sd, "pen_lifts_dense.0", x)).squeeze(-1)
    return eps, pen


# --------------------------------------------------------------------------
# diffusion schedule + chain
# --------------------------------------------------------------------------
def beta_schedule():
    """utils/nn.py:19-39, fp32 and in the reference's op order."""
    lin = torch.linspace(math.log(1e-5), math.log(0.4), NUM_DIFFUSION_STEPS)
    return 0.02 + torch.exp(lin)


def alpha_bar(beta):
    return torch.cumprod(1 - beta, dim=0)                         # inference.py:81


def stroke_length(n_tokens):
    """inference.py:77-78."""
    t = n_tokens * 16
    return t - (t % 8) + 8


def posterior_new(x, eps, beta, alpha, alpha_next, z):
    """utils/nn.py:110-112 with the noise draw z injected."""
    out = (x - torch.sqrt(1 - alpha) * eps) / torch.sqrt(1 - beta)
    return out + z * torch.sqrt(1 - alpha_next)


def posterior_standard(x, eps, beta, alpha, z, add_sigma):
    """utils/nn.py:84-87 with the noise draw z injected."""
    out = (1 / torch.sqrt(1 - beta)) * (x - (beta * eps / torch.sqrt(1 - alpha)))
    if add_sigma:
        out = out + torch.sqrt(beta) * z
    return out


@torch.no_grad()
def reverse_chain(sd, text_ids, style, x0, noise, diffusion_mode="new", steps=None):
    """inference.py:81-96 with x0 and the per-step noise injected.

    noise: [60,B,T,2]; noise[i] is the draw consumed at loop index i (the loop
    runs i = 59..0).  Returns strokes [B,T,3] = cat(x, pen of the last step).
    `steps` (default all 60) limits the loop to the first `steps` iterations --
    used by bounded CPU-baseline timing only."""
    dt = x0.dtype
    beta = beta_schedule()
    abar = alpha_bar(beta)
    B = x0.shape[0]
    x = x0.clone()
    pen = None
    order = list(range(NUM_DIFFUSION_STEPS - 1, -1, -1))
    if steps is not None:
        order = order[:steps]
    for i in order:
        a = (abar[i] * torch.ones((B, 1, 1))).to(dt)
        b = (beta[i] * torch.ones((B, 1, 1))).to(dt)
        a_next = (abar[i - 1] if i > 1 else torch.tensor(1.0)).to(dt)
        eps, pen = denoiser_forward(sd, x, text_ids, torch.sqrt(a), style)
        if diffusion_mode == "standard":
            x = posterior_standard(x, eps, b, a, noise[i].to(dt), add_sigma=bool(i))
        else:
            x = posterior_new(x, eps, b, a, a_next, noise[i].to(dt))
    return torch.cat((x, pen.unsqueeze(2)), dim=2)
