"""Plain fp32 PyTorch reference of the fused tcgen05 GEMM + row epilogue (csrc/common.cuh order of
operations), and a ctypes driver for the `dhg_debug_tc_gemm_ex` test hook.  Shared by the GPU tests
and tools/gemm_bench.py."""
import ctypes

import torch

from dhg_b200 import _abi


def split_pack(x):
    """fp32 [rows, C] (C % 32 == 0) -> bf16 [rows, 2C] in split storage (csrc/common.cuh `bfs`): per group of 32
    elements the 32 hi halves, then the 32 lo halves; value = hi + lo."""
    x = x.float()
    hi = x.bfloat16()
    lo = (x - hi.float()).bfloat16()
    r, c = x.shape
    return torch.stack((hi.view(r, c // 32, 32), lo.view(r, c // 32, 32)), dim=2).reshape(r, 2 * c).contiguous()


def split_unpack(w):
    r, c2 = w.shape
    g = w.float().view(r, c2 // 64, 2, 32)
    return (g[:, :, 0] + g[:, :, 1]).reshape(r, c2 // 2)


def split_weights(w):
    """[taps, N, K] fp32 -> [taps, N, 2K] bf16 with K grouped like the activations (w_hi x 32 | w_lo x 32)."""
    t, n, k = w.shape
    return split_pack(w.reshape(t * n, k)).view(t, n, 2 * k).contiguous()


def make_split_case(rows, K, N, taps, **kw):
    """The same case in split I/O (fp32-contract mode): fp32 values, every activation operand stored as bfs pairs."""
    device = kw.get("device", "cuda")
    c = make_case(rows, K, N, taps, **kw)
    g = torch.Generator().manual_seed(kw.get("seed", 0) + 99)
    f = {}   # the fp32 values behind every operand (the reference works on these)
    a = torch.randn(rows, K, generator=g)
    a[c["pad"].cpu()] = 0
    f["a"] = a.to(device)
    f["w"] = (torch.randn(taps, N, K, generator=g) / (K * taps) ** 0.5).to(device)
    for k in ("rowbias", "res_pre", "res_post"):
        f[k] = torch.randn(c[k].shape, generator=g).to(device) if c[k] is not None else None
    c["f32"] = f
    c["a"], c["w"] = split_pack(f["a"]), split_weights(f["w"])
    for k in ("rowbias", "res_pre", "res_post"):
        c[k] = split_pack(f[k]) if f[k] is not None else None
    for k in ("out_raw", "out_act"):
        if c[k] is not None:
            c[k] = torch.full((rows, 2 * N), float("nan"), dtype=torch.bfloat16, device=device)
    c["split"] = True
    return c


def make_case(rows, K, N, taps, *, period=None, pad_first=0, bias=True, rowbias=False, res_pre=False, ln=False,
              film=0, res_post=False, up=False, raw=True, act=False, dot=0, seed=0, device="cuda", dual_K2=0, variants=3):
    """film: 0 none, 1 one vector for the batch (bstride 0), 2 per-sample vectors.
    dot: 0 normal stores; 1 / 2 dot mode on the value / on SiLU(value) (no stored outputs)."""
    g = torch.Generator().manual_seed(seed)
    period = period or rows
    nb = (rows + period - 1) // period
    nvalid = (rows // period) * period if period != rows else rows
    c = {"rows": rows, "K": K, "N": N, "taps": taps, "period": period, "pad_first": pad_first, "nvalid": nvalid, "ln": ln}
    r = torch.arange(rows)
    pad = (r >= nvalid) | ((r % period == 0) if pad_first else torch.zeros(rows, dtype=torch.bool))
    a = torch.randn(rows, K, generator=g)
    a[pad] = 0
    c["a"] = a.bfloat16().to(device)
    c["w"] = (torch.randn(taps, N, K, generator=g) / (K * taps) ** 0.5).bfloat16().to(device)
    c["pad"] = pad.to(device)
    c["bias"] = torch.randn(N, generator=g).to(device) if bias else None
    # per-position term (bf16) for the first rowbias_cols columns only (q / k get positional embeddings, v does not)
    c["rowbias_cols"] = (N if N <= 256 else (2 * N // 3) // 32 * 32) if rowbias else 0
    c["rowbias"] = torch.randn(period - pad_first, c["rowbias_cols"], generator=g).bfloat16().to(device) if rowbias else None
    c["res_pre"] = torch.randn(rows, N, generator=g).bfloat16().to(device) if res_pre else None
    if film == 1:
        c["gamma"], c["beta"], c["bstride"] = (1 + 0.3 * torch.randn(N, generator=g)).to(device), torch.randn(N, generator=g).to(device), 0
    elif film == 2:
        c["gamma"] = (1 + 0.3 * torch.randn(nb, 2 * N, generator=g)).to(device)
        c["beta"], c["bstride"] = c["gamma"][:, N:], 2 * N
    else:
        c["gamma"] = c["beta"] = None
        c["bstride"] = 0
    c["up"] = up
    if res_post and up:
        plo = (period - pad_first) // 2 + 1
        c["period_lo"] = plo
        c["res_post"] = torch.randn(nb * plo + 1, N, generator=g).bfloat16().to(device)
    elif res_post:
        c["period_lo"] = 0
        c["res_post"] = torch.randn(rows, N, generator=g).bfloat16().to(device)
    else:
        c["period_lo"], c["res_post"] = 0, None
    c["out_raw"] = torch.full((rows, N), float("nan"), dtype=torch.bfloat16, device=device) if raw and not dot else None
    c["out_act"] = torch.full((rows, N), float("nan"), dtype=torch.bfloat16, device=device) if act and not dot else None
    c["dot_act"] = 1 if dot == 2 else 0
    c["dot_w"] = (torch.randn(3, N, generator=g) / N ** 0.5).to(device) if dot else None
    c["dot_out"] = torch.full((rows, 4), float("nan"), device=device) if dot else None
    # dual-operand mode: a second A matrix against 3-tap weights joins the accumulation; the first operand's weights
    # are one of `variants` stacked [N, K] matrices (engine.cu: fc with the step's FiLM scale folded in)
    c["dual_K2"] = dual_K2
    if dual_K2:
        a2 = torch.randn(rows, dual_K2, generator=g)
        a2[pad] = 0
        c["a2"] = a2.bfloat16().to(device)
        c["w2"] = (torch.randn(3, N, dual_K2, generator=g) / (3 * dual_K2) ** 0.5).bfloat16().to(device)
        c["w"] = (torch.randn(variants, N, K, generator=g) / K ** 0.5).bfloat16().to(device)
        c["variant"] = variants - 1
    return c


def reference(c):
    rows, N, taps, period, pf = c["rows"], c["N"], c["taps"], c["period"], c["pad_first"]
    if c.get("split"):   # fp64 reference on the fp32 values themselves
        c = dict(c, **{k: (v.double() if v is not None else None) for k, v in c["f32"].items()})
        c = dict(c, bias=c["bias"].double() if c["bias"] is not None else None,
                 gamma=c["gamma"].double() if c["gamma"] is not None else None, beta=c["beta"].double() if c["beta"] is not None else None)
        af, wf = c["a"], c["w"]
    else:
        af, wf = c["a"].float(), c["w"].float()
    x = torch.zeros(rows, N, device=af.device, dtype=af.dtype)
    if c.get("dual_K2"):
        wf = wf[c["variant"]:c["variant"] + 1]
        a2, w2 = c["a2"].float(), c["w2"].float()
        for t in range(3):
            shift = t - 1
            src = torch.zeros_like(a2)
            lo, hi = max(0, -shift), min(rows, rows - shift)
            src[lo:hi] = a2[lo + shift:hi + shift]
            x += src @ w2[t].T
    for t in range(taps):
        shift = t - taps // 2
        src = torch.zeros_like(af)
        lo, hi = max(0, -shift), min(rows, rows - shift)
        src[lo:hi] = af[lo + shift:hi + shift]
        x += src @ wf[t].T
    r = torch.arange(rows, device=af.device)
    b = r // period
    pos = (r % period - pf).clamp(min=0)
    if c["bias"] is not None:
        x += c["bias"][None]
    if c["rowbias"] is not None:
        x[:, :c["rowbias_cols"]] += c["rowbias"].to(x.dtype)[pos]
    if c["res_pre"] is not None:
        x += c["res_pre"].to(x.dtype)
    if c["ln"]:
        x = torch.nn.functional.layer_norm(x, (N,), eps=1e-6)
    if c["gamma"] is not None:
        if c["bstride"]:
            x = x * c["gamma"][b, :N] + c["gamma"][b, N:]
        else:
            x = x * c["gamma"][None] + c["beta"][None]
    if c["res_post"] is not None:
        if c["up"]:
            x += c["res_post"].to(x.dtype)[b * c["period_lo"] + 1 + pos // 2]
        else:
            x += c["res_post"].to(x.dtype)
    x[c["pad"]] = 0
    return x


def run(lib, c, repeats=0, allow_unavailable=False):
    """allow_unavailable: return None (instead of failing) when the kernel has no plan for the forced tile
    configuration (tune_* options) on this shape."""
    p = lambda t: ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p(0)
    N = c["N"]
    e = _abi.DebugEpilogue(
        p(c["bias"]), p(c["rowbias"]), c["rowbias_cols"], p(c["res_pre"]), N, int(c["ln"]), p(c["gamma"]), p(c["beta"]), c["bstride"],
        p(c["res_post"]), N, int(c["up"]), c["period_lo"], p(c["out_raw"]), N, p(c["out_act"]), N,
        c["period"], c["pad_first"], c["nvalid"], p(c.get("dot_w")), p(c.get("dot_out")), int(c.get("dot_act", 0)),
        1 if c.get("split") else 0,
        p(c.get("a2")), c.get("dual_K2", 0), c.get("dual_K2", 0), p(c.get("w2")), (c["w"].shape[0] * N) if c.get("dual_K2") else 0,
        (c["variant"] * N) if c.get("dual_K2") else 0)
    ms = ctypes.c_float(0)
    km = 2 if c.get("split") else 1   # split I/O: K and lda in bf16 units
    rc = lib.dhg_debug_tc_gemm_ex(0, p(c["a"]), km * c["K"], c["rows"], p(c["w"]), km * c["K"], N, c["taps"], ctypes.byref(e),
                                  repeats, ctypes.byref(ms), ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
    if rc != 0 and allow_unavailable:
        msg = lib.dhg_last_error().decode()
        assert any(k in msg for k in ("does not fit", "do not fit", "not enough shared memory", "too many column groups", "does not fit",
                                      "A ring too small")), msg
        return None
    assert rc == 0, lib.dhg_last_error().decode()
    torch.cuda.synchronize()
    return ms.value
