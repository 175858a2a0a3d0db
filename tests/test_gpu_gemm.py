"""The tcgen05/TMA GEMM kernel on its own, through the C-ABI test hook, against a plain fp32
PyTorch reference of the same op (bf16 inputs, fp32 accumulate), over every (K, N, taps) family the
denoiser uses, ragged row counts, the K=96 tail and the 3-tap row-shift with zero halo."""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

import pytest
import torch

pytestmark = pytest.mark.gpu

CASES = [
    # rows, K, N, taps
    (1000, 128, 128, 1), (257, 384, 192, 1), (300, 192, 384, 1), (513, 128, 64, 3), (400, 96, 192, 3),
    (130, 64, 128, 3), (1000, 256, 768, 1), (640, 384, 1152, 1), (200, 768, 384, 1), (50, 128, 96, 3),
    (4096, 384, 256, 3), (127, 192, 576, 1), (128, 256, 512, 1), (2000, 192, 128, 3),
]


@pytest.mark.parametrize("rows,K,N,taps", CASES)
def test_tc_gemm_matches_torch(built_lib, rows, K, N, taps):
    g = torch.Generator().manual_seed(rows * 7 + K + N + taps)
    a = torch.randn(rows, K, generator=g).bfloat16().cuda()
    w = (torch.randn(taps, N, K, generator=g) / K ** 0.5).bfloat16().cuda()
    bias = torch.randn(N, generator=g).cuda()
    out = torch.full((rows, N), float("nan"), dtype=torch.bfloat16, device="cuda")
    p = lambda t: ctypes.c_void_p(t.data_ptr())
    rc = built_lib.dhg_debug_tc_gemm(0, p(a), K, rows, p(w), K, N, taps, p(bias), p(out),
                                     ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
    assert rc == 0, built_lib.dhg_last_error().decode()
    torch.cuda.synchronize()
    af, wf = a.float(), w.float()
    ref = bias[None, :].repeat(rows, 1)
    for t in range(taps):
        shift = t - taps // 2
        src = torch.zeros_like(af)
        lo, hi = max(0, -shift), min(rows, rows - shift)
        src[lo:hi] = af[lo + shift:hi + shift]
        ref += src @ wf[t].T
    err = (out.float() - ref).abs().max().item()
    assert torch.isfinite(out.float()).all()
    assert err < 2e-2 * max(1.0, ref.abs().max().item()), err


# every fused-epilogue combination the denoiser uses (engine.cu: conv_block / encoder_layer / build_plan)
EPI_CASES = [
    # name, rows, K, N, taps, kwargs
    ("conv_film_act", 1000, 128, 64, 3, dict(period=50, pad_first=1, film=1, raw=False, act=True)),
    ("conv_film_act_n96", 700, 128, 96, 3, dict(period=99, pad_first=1, film=1, raw=False, act=True)),
    ("conv_skip_raw", 1500, 192, 128, 3, dict(period=197, pad_first=1)),
    ("fc_film_respost", 1300, 128, 128, 1, dict(period=99, pad_first=1, film=1, res_post=True, act=True)),
    ("fc_film_respost_256", 900, 256, 256, 1, dict(period=99, pad_first=1, film=1, res_post=True)),
    ("skipconv_up", 1183, 128, 192, 3, dict(period=169, pad_first=1, res_post=True, up=True, act=True)),
    ("skipconv_up_384", 797, 256, 384, 3, dict(period=99, pad_first=1, res_post=True, up=True, act=True)),
    ("rowbias_q", 1379, 192, 192, 1, dict(period=197, pad_first=1, rowbias=True)),
    ("rowbias_qkv", 1379, 192, 576, 1, dict(period=197, pad_first=1, rowbias=True)),
    ("rowbias_qkv_1152", 451, 384, 1152, 1, dict(period=50, pad_first=1, rowbias=True)),
    ("rowbias_text", 240, 256, 512, 1, dict(period=24, pad_first=0, rowbias=True)),
    ("ln_film_respost_192", 1379, 192, 192, 1, dict(period=197, pad_first=1, ln=True, film=1, res_post=True)),
    ("ln_film_respre_256", 991, 256, 256, 1, dict(period=99, pad_first=1, ln=True, film=1, res_pre=True, act=True)),
    ("ln_film_respre_384", 451, 384, 384, 1, dict(period=50, pad_first=1, ln=True, film=1, res_pre=True)),
    ("ln_film_respost_384", 451, 768, 384, 1, dict(period=50, pad_first=1, ln=True, film=1, res_post=True)),
    ("ln_film_text", 240, 384, 192, 1, dict(period=24, pad_first=0, ln=True, film=1)),
    ("ln_only_style", 700, 768, 384, 1, dict(period=70, pad_first=0, ln=True)),
    ("film_per_sample", 1000, 128, 128, 3, dict(period=50, pad_first=1, film=2, raw=False, act=True)),
    ("ln_film_per_sample", 451, 384, 384, 1, dict(period=50, pad_first=1, ln=True, film=2, res_pre=True)),
    ("ffn1_act_768", 451, 384, 768, 1, dict(period=50, pad_first=1, raw=False, act=True)),
    ("many_tiles", 40000, 128, 128, 3, dict(period=393, pad_first=1, film=1, raw=False, act=True)),
    ("many_tiles_ln384", 30000, 384, 384, 1, dict(period=50, pad_first=1, ln=True, film=1, res_pre=True)),
]


@pytest.mark.parametrize("name,rows,K,N,taps,kw", EPI_CASES, ids=[c[0] for c in EPI_CASES])
def test_tc_gemm_fused_epilogue(built_lib, name, rows, K, N, taps, kw):
    import gemm_ref

    c = gemm_ref.make_case(rows, K, N, taps, seed=len(name) + rows, **kw)
    gemm_ref.run(built_lib, c)
    ref = gemm_ref.reference(c)
    scale = max(1.0, ref.abs().max().item())
    if c["out_raw"] is not None:
        assert torch.isfinite(c["out_raw"].float()).all()
        err = (c["out_raw"].float() - ref).abs().max().item()
        assert err < 2e-2 * scale, (name, "raw", err)
    if c["out_act"] is not None:
        sref = torch.nn.functional.silu(ref)
        assert torch.isfinite(c["out_act"].float()).all()
        err = (c["out_act"].float() - sref).abs().max().item()
        assert err < 2e-2 * scale, (name, "act", err)


SPLIT_CASES = [c for c in EPI_CASES if c[0] in (
    "conv_film_act", "conv_film_act_n96", "conv_skip_raw", "fc_film_respost", "skipconv_up", "skipconv_up_384", "rowbias_q", "rowbias_qkv_1152",
    "rowbias_text", "ln_film_respost_192", "ln_film_respre_256", "ln_film_respre_384", "ln_film_respost_384", "ln_only_style",
    "film_per_sample", "ln_film_per_sample", "ffn1_act_768", "many_tiles", "many_tiles_ln384")]


@pytest.mark.parametrize("name,rows,K,N,taps,kw", SPLIT_CASES, ids=[c[0] for c in SPLIT_CASES])
def test_tc_gemm_split_io_meets_fp32_contract(built_lib, name, rows, K, N, taps, kw):
    """Split I/O (the fp32-contract mode, DHG_PREC_FP32 on tcgen05): operands are {bf16 hi, bf16 lo} pairs, the GEMM is
    (hi + lo) . w_hi + hi . w_lo in bf16 MMAs.  Against an fp64 reference on the fp32 values: 1e-4 of the output scale
    (the products carry ~2^-17 each, the stored result another 2^-17), three orders below the bf16 mode."""
    import gemm_ref

    c = gemm_ref.make_split_case(rows, K, N, taps, seed=len(name) + rows, **kw)
    gemm_ref.run(built_lib, c)
    ref = gemm_ref.reference(c)
    scale = max(1.0, ref.abs().max().item())
    if c["out_raw"] is not None:
        got = gemm_ref.split_unpack(c["out_raw"]).double()
        assert torch.isfinite(got).all()
        err = (got - ref).abs().max().item()
        assert err < 1e-4 * scale, (name, "raw", err)
    if c["out_act"] is not None:
        got = gemm_ref.split_unpack(c["out_act"]).double()
        assert torch.isfinite(got).all()
        err = (got - torch.nn.functional.silu(ref)).abs().max().item()
        assert err < 1e-4 * scale, (name, "act", err)


DUAL_CASES = [
    # the five ConvBlocks whose conv_skip is contracted inside their last GEMM (engine.cu skip fusion): K = Cout, K2 = Cin
    ("dual_enc1", 3000, 128, 128, 1, dict(period=393, pad_first=1, dual_K2=128)),
    ("dual_enc2", 1971, 192, 192, 1, dict(period=197, pad_first=1, dual_K2=128)),
    ("dual_enc4", 991, 256, 256, 1, dict(period=99, pad_first=1, dual_K2=192)),
    ("dual_dec3", 991, 256, 256, 1, dict(period=99, pad_first=1, dual_K2=384, act=True)),
    ("dual_dec2", 1971, 192, 192, 1, dict(period=197, pad_first=1, dual_K2=256)),
    ("dual_many_tiles", 40000, 128, 128, 1, dict(period=393, pad_first=1, dual_K2=128)),
    ("dual_many_tiles_192", 30000, 192, 192, 1, dict(period=197, pad_first=1, dual_K2=256)),
]


@pytest.mark.parametrize("name,rows,K,N,taps,kw", DUAL_CASES, ids=[c[0] for c in DUAL_CASES])
def test_tc_gemm_dual_operand(built_lib, name, rows, K, N, taps, kw):
    """Dual-operand mode: a2 . W1[variant]^T + conv3(x, W2) + bias in one accumulation (two A matrices, two tensor maps)."""
    import gemm_ref

    c = gemm_ref.make_case(rows, K, N, taps, seed=len(name) + rows, **kw)
    gemm_ref.run(built_lib, c)
    ref = gemm_ref.reference(c)
    scale = max(1.0, ref.abs().max().item())
    assert torch.isfinite(c["out_raw"].float()).all()
    assert (c["out_raw"].float() - ref).abs().max().item() < 2e-2 * scale, name
    if c["out_act"] is not None:
        assert (c["out_act"].float() - torch.nn.functional.silu(ref)).abs().max().item() < 2e-2 * scale, name


DOT_CASES = [
    ("dot_conv_skip", 3000, 192, 128, 3, dict(period=393, pad_first=1, dot=1)),
    ("dot_conv2_film_act", 3000, 64, 128, 3, dict(period=393, pad_first=1, film=1, dot=2)),
    ("dot_many_tiles", 40000, 64, 128, 3, dict(period=393, pad_first=1, film=1, dot=2)),
    ("dot_n256", 2000, 128, 256, 1, dict(period=99, pad_first=1, dot=1)),
    ("dot_n32_folded_conv", 40000, 192, 32, 3, dict(period=393, pad_first=1, dot=1)),
]


@pytest.mark.parametrize("name,rows,K,N,taps,kw", DOT_CASES, ids=[c[0] for c in DOT_CASES])
def test_tc_gemm_dot_mode(built_lib, name, rows, K, N, taps, kw):
    """Dot mode (engine.cu tail fusion): the row is not stored, only its 3 dot products with fp32 vectors."""
    import gemm_ref

    c = gemm_ref.make_case(rows, K, N, taps, seed=len(name) + rows, **kw)
    gemm_ref.run(built_lib, c)
    x = gemm_ref.reference(c)
    if c["dot_act"]:
        x = torch.nn.functional.silu(x)
    ref = x @ c["dot_w"].T
    got = c["dot_out"][:, :3]
    assert torch.isfinite(got).all()
    err = (got - ref).abs().max().item()
    assert err < 1e-2 * max(1.0, ref.abs().max().item()), (name, err)   # bf16 inputs, fp32 accumulation, tanh.approx SiLU


# The plan-time autotuner (engine.cu Builder::autotune) may pick any of these for a GEMM of the denoiser, so every one
# must compute the same bits as the built-in rule: tile width x interleaved accumulators x {resident W, streamed W,
# streamed W + CTA pairs}.
TUNE_CASES = [
    ("conv_film_act", 20000, 128, 64, 3, dict(period=50, pad_first=1, film=1, raw=False, act=True)),
    ("fc_film_respost", 21000, 128, 128, 1, dict(period=99, pad_first=1, film=1, res_post=True, act=True)),
    ("skipconv_up", 23660, 128, 192, 3, dict(period=169, pad_first=1, res_post=True, up=True, act=True)),
    ("rowbias_qkv", 20685, 192, 576, 1, dict(period=197, pad_first=1, rowbias=True)),
    ("rowbias_qkv_1152", 20050, 384, 1152, 1, dict(period=50, pad_first=1, rowbias=True)),
    ("ffn1_act_768", 20050, 384, 768, 1, dict(period=50, pad_first=1, raw=False, act=True)),
    ("conv_skip_256", 19999, 384, 256, 3, dict(period=99, pad_first=1)),
    ("ln_film_respre_384", 20050, 384, 384, 1, dict(period=50, pad_first=1, ln=True, film=1, res_pre=True)),
    ("ln_film_respost_192", 20685, 192, 192, 1, dict(period=197, pad_first=1, ln=True, film=1, res_post=True)),
    ("ln_film_respre_256_act", 19999, 512, 256, 1, dict(period=99, pad_first=1, ln=True, film=1, res_pre=True, act=True)),
    ("ln_film_k768_384_act", 20050, 768, 384, 1, dict(period=50, pad_first=1, ln=True, film=1, raw=False, act=True)),
    ("dual_fc_conv_skip_192", 20685, 192, 192, 1, dict(period=197, pad_first=1, dual_K2=256)),
    ("dual_fc_conv_skip_128", 19999, 128, 128, 1, dict(period=393, pad_first=1, dual_K2=128)),
    ("tiny_rows_qkv", 3, 384, 1152, 1, dict(period=2, pad_first=1, rowbias=True)),
    ("tiny_rows_ln", 3, 768, 384, 1, dict(period=2, pad_first=1, ln=True, film=1, res_pre=True)),
]


@pytest.mark.parametrize("name,rows,K,N,taps,kw", TUNE_CASES, ids=[c[0] for c in TUNE_CASES])
def test_tc_gemm_every_tile_configuration_gives_the_same_bits(built_lib, name, rows, K, N, taps, kw):
    import gemm_ref

    def setopt(**o):
        for k, v in o.items():
            assert built_lib.dhg_set_option(None, f"tune_{k}".encode(), v) == 0

    def outputs(c):
        return [t.clone() for t in (c["out_raw"], c["out_act"]) if t is not None]

    try:
        setopt(bn=-1, g=-1, resident=-1, pair=-1)
        c = gemm_ref.make_case(rows, K, N, taps, seed=len(name) + rows, **kw)
        gemm_ref.run(built_lib, c)
        base = outputs(c)
        ref = gemm_ref.reference(c)
        first = base[0].float() if c["out_raw"] is not None else None
        if first is not None:
            assert (first - ref).abs().max().item() < 2e-2 * max(1.0, ref.abs().max().item())
        tried = 0
        bns = [N] if kw.get("ln") else [b for b in (384, 256, 192, 128, 96, 64) if N % b == 0]
        for bn in bns:
            for g in (1, 2, 4):
                if g > 1 and g * bn > 256:
                    continue
                for resident, pair in ((1, 0), (0, 0), (0, 1), (1, 2), (0, 2)):   # pair 2: column-split LayerNorm cluster
                    if pair and g != 1:
                        continue
                    if pair == 2 and not (kw.get("ln") and N % 128 == 0):
                        continue
                    setopt(bn=-1 if kw.get("ln") else bn, g=g, resident=resident, pair=pair)
                    for t in (c["out_raw"], c["out_act"]):
                        if t is not None:
                            t.fill_(float("nan"))
                    if gemm_ref.run(built_lib, c, allow_unavailable=True) is None:
                        continue
                    tried += 1
                    for got, want in zip(outputs(c), base):
                        assert torch.equal(got.view(torch.int16), want.view(torch.int16)), (name, bn, g, resident, pair)
        assert tried >= 2, tried
        # row tiles walked from the last to the first (the engine alternates the direction from kernel to kernel)
        setopt(bn=-1, g=-1, resident=-1, pair=-1, rev=1)
        gemm_ref.run(built_lib, c)
        for got, want in zip(outputs(c), base):
            assert torch.equal(got.view(torch.int16), want.view(torch.int16)), (name, "reverse")
    finally:
        setopt(bn=-1, g=-1, resident=-1, pair=-1, rev=0)


SPLIT_TUNE_CASES = [
    ("conv_film_act", 20000, 128, 64, 3, dict(period=50, pad_first=1, film=1, raw=False, act=True)),
    ("rowbias_qkv", 20685, 192, 576, 1, dict(period=197, pad_first=1, rowbias=True)),
    ("ffn1_act_768", 20050, 384, 768, 1, dict(period=50, pad_first=1, raw=False, act=True)),
    ("ln_film_respre_384", 20050, 384, 384, 1, dict(period=50, pad_first=1, ln=True, film=1, res_pre=True)),
    ("skipconv_up", 23660, 128, 192, 3, dict(period=169, pad_first=1, res_post=True, up=True, act=True)),
    ("tiny_rows_qkv", 3, 384, 1152, 1, dict(period=2, pad_first=1, rowbias=True)),
    ("tiny_rows_conv", 10, 128, 128, 3, dict(period=9, pad_first=1, film=1, raw=False, act=True)),
    ("tiny_rows_ln", 3, 768, 384, 1, dict(period=2, pad_first=1, ln=True, film=1, res_pre=True)),
    ("tiny_enc2_conv2", 6, 96, 192, 3, dict(period=5, pad_first=1, film=2, raw=False, act=True)),
    ("tiny_enc2_conv1", 6, 128, 96, 3, dict(period=5, pad_first=1, film=2, raw=False, act=True)),
    ("small_k96", 700, 96, 192, 3, dict(period=99, pad_first=1, film=1, raw=False, act=True)),
]


@pytest.mark.parametrize("name,rows,K,N,taps,kw", SPLIT_TUNE_CASES, ids=[c[0] for c in SPLIT_TUNE_CASES])
def test_tc_gemm_split_io_every_tile_configuration_gives_the_same_bits(built_lib, name, rows, K, N, taps, kw):
    """The plan-time tuner picks among these by timing, so every one of them must be right (and bit-identical) in split
    I/O too, down to matrices of a few rows (T = 8 -> one stroke row per sample at the deepest level)."""
    import gemm_ref

    def setopt(**o):
        for k, v in o.items():
            assert built_lib.dhg_set_option(None, f"tune_{k}".encode(), v) == 0

    def outputs(c):
        return [t.clone() for t in (c["out_raw"], c["out_act"]) if t is not None]

    try:
        setopt(bn=-1, g=-1, resident=-1, pair=-1)
        c = gemm_ref.make_split_case(rows, K, N, taps, seed=len(name) + rows, **kw)
        gemm_ref.run(built_lib, c)
        base = outputs(c)
        ref = gemm_ref.reference(c)
        got = gemm_ref.split_unpack(base[0]).double()
        want = ref if c["out_raw"] is not None else torch.nn.functional.silu(ref)
        assert (got - want).abs().max().item() < 1e-4 * max(1.0, want.abs().max().item())
        tried = 0
        bns = [N] if kw.get("ln") else [b for b in (384, 256, 192, 128, 96, 64) if N % b == 0]
        for bn in bns:
            for resident, pair in ((1, 0), (0, 0), (0, 1)):
                setopt(bn=-1 if kw.get("ln") else bn, g=1, resident=resident, pair=pair)
                for t in (c["out_raw"], c["out_act"]):
                    if t is not None:
                        t.fill_(float("nan"))
                if gemm_ref.run(built_lib, c, allow_unavailable=True) is None:
                    continue
                tried += 1
                for g_, w_ in zip(outputs(c), base):
                    assert torch.equal(g_.view(torch.int16), w_.view(torch.int16)), (name, bn, resident, pair)
        assert tried >= 1, tried
        setopt(bn=-1, g=-1, resident=-1, pair=-1, rev=1)
        gemm_ref.run(built_lib, c)
        for g_, w_ in zip(outputs(c), base):
            assert torch.equal(g_.view(torch.int16), w_.view(torch.int16)), (name, "reverse")
    finally:
        setopt(bn=-1, g=-1, resident=-1, pair=-1, rev=0)
