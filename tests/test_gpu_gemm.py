"""The tcgen05/TMA GEMM kernel on its own, through the C-ABI test hook, against a plain fp32
PyTorch reference of the same op (bf16 inputs, fp32 accumulate), over every (K, N, taps) family the
denoiser uses, ragged row counts, the K=96 tail and the 3-tap row-shift with zero halo."""
import ctypes

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
