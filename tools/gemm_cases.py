"""The GEMM launches of one denoiser step at batch B (shapes and fused epilogues as planned by csrc/engine.cu),
and a runner that times each through the C-ABI test hook.  Used by tools/gemm_bench.py and bench.py."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "diffusion-handwriting-generation.pytorch_b200"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def step_gemm_cases(B, T=392, L=24, SP=70, chain=True):
    """[(name, launches per step, rows, K, N, taps, epilogue kwargs)]

    chain=True: the GEMMs a step of the sampling chain launches (59): dec1.fc is folded into the head kernel
     and in the other five ConvBlocks conv_skip is contracted inside the block's last GEMM (dual-operand
    launches); chain=False: all 65 of a stand-alone forward (dhg_denoise)."""
    lv = [dict(period=(T >> l) + 1, pad_first=1) for l in range(4)]
    R = [B * ((T >> l) + 1) + 1 for l in range(4)]
    tx, stl = dict(period=L, pad_first=0), dict(period=SP, pad_first=0)
    RT, RS = B * L, B * SP
    CASES = [
        # name, count per step, rows, K, N, taps, kwargs
        ("L0 conv_skip 128->128", 0 if chain else 1, R[0], 128, 128, 3, dict(**lv[0])),
        ("L0 conv1 128->64 film act", 1, R[0], 128, 64, 3, dict(**lv[0], film=1, raw=False, act=True)),
        ("L0 conv2 64->128 film act", 2, R[0], 64, 128, 3, dict(**lv[0], film=1, raw=False, act=True)),
        ("L0 fc 128 film +skip", 0 if chain else 2, R[0], 128, 128, 1, dict(**lv[0], film=1, res_post=True)),
        ("L0 enc1 fc+conv_skip dual", 1 if chain else 0, R[0], 128, 128, 1, dict(**lv[0], dual_K2=128)),
        ("L0 skip_conv1 128->192 up", 1, R[0], 128, 192, 3, dict(**lv[0], res_post=True, up=True, act=True)),
        ("L0 dec1.conv_skip 192->128", 1, R[0], 192, 128, 3, dict(**lv[0])),
        ("L0 dec1.conv1 192->64", 1, R[0], 192, 64, 3, dict(**lv[0], film=1, raw=False, act=True)),
        ("L1 enc2.conv_skip 128->192", 0 if chain else 1, R[1], 128, 192, 3, dict(**lv[1])),
        ("L1 enc2 fc+conv_skip dual", 1 if chain else 0, R[1], 192, 192, 1, dict(**lv[1], dual_K2=128)),
        ("L1 dec2 fc+conv_skip dual", 1 if chain else 0, R[1], 192, 192, 1, dict(**lv[1], dual_K2=256)),
        ("L1 conv1 128->96", 1, R[1], 128, 96, 3, dict(**lv[1], film=1, raw=False, act=True)),
        ("L1 conv2 96->192", 2, R[1], 96, 192, 3, dict(**lv[1], film=1, raw=False, act=True)),
        ("L1 fc 192 +skip", 0 if chain else 2, R[1], 192, 192, 1, dict(**lv[1], film=1, res_post=True)),
        ("L1 wq 192 rowbias", 1, R[1], 192, 192, 1, dict(**lv[1], rowbias=True)),
        ("L1 dense LN film +x", 1, R[1], 192, 192, 1, dict(**lv[1], ln=True, film=1, res_post=True)),
        ("L1 qkv 192->576 rowbias", 1, R[1], 192, 576, 1, dict(**lv[1], rowbias=True)),
        ("L1 dense2 res LN film", 1, R[1], 192, 192, 1, dict(**lv[1], ln=True, film=1, res_pre=True, act=True)),
        ("L1 ffn1 192->384 act", 1, R[1], 192, 384, 1, dict(**lv[1], raw=False, act=True)),
        ("L1 ffn3 384->192 res LN film", 1, R[1], 384, 192, 1, dict(**lv[1], ln=True, film=1, res_pre=True)),
        ("L1 skip_conv2 192->256 up", 1, R[1], 192, 256, 3, dict(**lv[1], res_post=True, up=True, act=True)),
        ("L1 dec2.conv_skip 256->192", 0 if chain else 1, R[1], 256, 192, 3, dict(**lv[1])),
        ("L1 dec2.conv1 256->96", 1, R[1], 256, 96, 3, dict(**lv[1], film=1, raw=False, act=True)),
        ("L2 enc4.conv_skip 192->256", 0 if chain else 1, R[2], 192, 256, 3, dict(**lv[2])),
        ("L2 enc4 fc+conv_skip dual", 1 if chain else 0, R[2], 256, 256, 1, dict(**lv[2], dual_K2=192)),
        ("L2 dec3 fc+conv_skip dual", 1 if chain else 0, R[2], 256, 256, 1, dict(**lv[2], dual_K2=384)),
        ("L2 conv1 192->128", 1, R[2], 192, 128, 3, dict(**lv[2], film=1, raw=False, act=True)),
        ("L2 conv2 128->256", 2, R[2], 128, 256, 3, dict(**lv[2], film=1, raw=False, act=True)),
        ("L2 fc 256 +skip", 0 if chain else 2, R[2], 256, 256, 1, dict(**lv[2], film=1, res_post=True)),
        ("L2 wq 256 rowbias", 1, R[2], 256, 256, 1, dict(**lv[2], rowbias=True)),
        ("L2 dense LN film +x", 1, R[2], 256, 256, 1, dict(**lv[2], ln=True, film=1, res_post=True)),
        ("L2 qkv 256->768", 1, R[2], 256, 768, 1, dict(**lv[2], rowbias=True)),
        ("L2 dense2 res LN film", 1, R[2], 256, 256, 1, dict(**lv[2], ln=True, film=1, res_pre=True, act=True)),
        ("L2 ffn1 256->512", 1, R[2], 256, 512, 1, dict(**lv[2], raw=False, act=True)),
        ("L2 ffn3 512->256", 1, R[2], 512, 256, 1, dict(**lv[2], ln=True, film=1, res_pre=True)),
        ("L2 skip_conv3 256->384 up", 1, R[2], 256, 384, 3, dict(**lv[2], res_post=True, up=True, act=True)),
        ("L2 dec3.conv_skip 384->256", 0 if chain else 1, R[2], 384, 256, 3, dict(**lv[2])),
        ("L2 dec3.conv1 384->128", 1, R[2], 384, 128, 3, dict(**lv[2], film=1, raw=False, act=True)),
        ("L3 att_dense 256->384", 1, R[3], 256, 384, 1, dict(**lv[3])),
        ("L3 wq 384 rowbias", 2, R[3], 384, 384, 1, dict(**lv[3], rowbias=True)),
        ("L3 dense LN film +x", 2, R[3], 384, 384, 1, dict(**lv[3], ln=True, film=1, res_post=True)),
        ("L3 qkv 384->1152", 2, R[3], 384, 1152, 1, dict(**lv[3], rowbias=True)),
        ("L3 dense2 res LN film", 2, R[3], 384, 384, 1, dict(**lv[3], ln=True, film=1, res_pre=True, act=True)),
        ("L3 ffn1 384->768", 2, R[3], 384, 768, 1, dict(**lv[3], raw=False, act=True)),
        ("L3 ffn3 768->384", 2, R[3], 768, 384, 1, dict(**lv[3], ln=True, film=1, res_pre=True)),
        ("TX text_dense 384->192 LN film", 1, RT, 384, 192, 1, dict(**tx, ln=True, film=1)),
        ("TX text_dense 384->256 LN film", 1, RT, 384, 256, 1, dict(**tx, ln=True, film=1)),
        ("TX text_dense 384->384 LN film", 2, RT, 384, 384, 1, dict(**tx, ln=True, film=1)),
        ("TX kv 192->384 rowbias", 1, RT, 192, 384, 1, dict(**tx, rowbias=True)),
        ("TX kv 256->512 rowbias", 1, RT, 256, 512, 1, dict(**tx, rowbias=True)),
        ("TX kv 384->768 rowbias", 2, RT, 384, 768, 1, dict(**tx, rowbias=True)),
        ("TS wq 384", 1, RT, 384, 384, 1, dict(**tx)),
        ("TS mha.dense res LN film act", 1, RT, 384, 384, 1, dict(**tx, ln=True, film=1, res_pre=True, raw=False, act=True)),
        ("TS text_ffn1 384->768 act", 1, RT, 384, 768, 1, dict(**tx, raw=False, act=True)),
        ("TS text_ffn3 768->384 LN film act", 1, RT, 768, 384, 1, dict(**tx, ln=True, film=1, raw=False, act=True)),
        ("ST style kv 384->768", 1, RS, 384, 768, 1, dict(**stl)),
    ]
    return CASES


def time_step_gemms(B, repeats=5, verbose=False):
    """Times every GEMM family; returns dict(us, flop, bytes) summed over one step's launches."""
    import gemm_ref
    from dhg_b200 import _abi

    lib = _abi.lib()
    tot_us = tot_fl = tot_by = 0.0
    rows_out = []
    for name, cnt, rows, K, N, taps, kw in step_gemm_cases(B):
        if cnt == 0:
            continue
        c = gemm_ref.make_case(rows, K, N, taps, seed=1, **kw)
        ms = gemm_ref.run(lib, c, repeats=repeats)
        K2 = kw.get("dual_K2", 0)
        fl = 2.0 * rows * N * (K * taps + 3 * K2)
        by = rows * (K + K2) * 2 + (taps * K + 3 * K2) * N * 2
        for k in ("out_raw", "out_act", "res_pre"):
            by += rows * N * 2 if c[k] is not None else 0
        if c["res_post"] is not None:
            by += c["res_post"].numel() * 2
        if c["rowbias"] is not None:
            by += c["rowbias"].numel() * 2
        us = ms * 1e3
        rows_out.append((name, cnt, us, fl, by))
        if verbose:
            print(f"      {name:38s} {us:8.1f} {fl/us/1e6:7.1f} {by/us/1e3:7.0f}  {cnt}", flush=True)
        tot_us += cnt * us; tot_fl += cnt * fl; tot_by += cnt * by
        del c
    return {"us": tot_us, "flop": tot_fl, "bytes": tot_by, "launches": sum(r[1] for r in rows_out), "rows": rows_out}
