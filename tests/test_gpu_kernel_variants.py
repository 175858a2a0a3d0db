"""The same GEMM unit cases under every kernel-selection switch (generic epilogue instance, forced CTA pairs
(cta_group::2), no interleaving / streaming weights): each variant runs in its own process because the switches
are read when the library is loaded (DHG_OPTS)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("opts", ["specialize=0", "pair=2", "pair=0,interleave=0,w_resident=0", "pdl=0",
                                  "direct_store=1,max_stages_a=2"])   # register -> global stores, shallow A ring
def test_gemm_cases_under_switch(opts):
    env = dict(os.environ, DHG_OPTS=opts)
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.join(ROOT, "tests", "test_gpu_gemm.py"), "-x", "-q", "-m", "gpu",
                        "-p", "no:cacheprovider"], env=env, capture_output=True, text=True, cwd=ROOT, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]


@pytest.mark.parametrize("opts", ["attn_early=0", "attn_max_slots=3"])
def test_attention_cases_with_late_tile_loads(opts):
    """attention_tc.cu requests the next item's Q/K/V right after P V by default; the plan-time tuner may switch a
    shape back to loading after O has been stored, so that order gets the same unit cases.  attn_max_slots=3: the
    kernel instances compiled for fewer slots (more registers per thread) on the shapes that normally run with 4-6."""
    env = dict(os.environ, DHG_OPTS=opts)
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.join(ROOT, "tests", "test_gpu_attention.py"), "-x", "-q", "-m", "gpu",
                        "-p", "no:cacheprovider"], env=env, capture_output=True, text=True, cwd=ROOT, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]


_CHAIN = r"""
import sys, numpy as np, torch
sys.path.insert(0, {root!r}); sys.path.insert(0, {pkg!r})
from dhg_b200 import DiffusionWriter
from oracle.dhg_oracle import init_state_dict
w = DiffusionWriter(state_dict=init_state_dict(0), num_layers=2, channels=128, dtype="bf16")
if len(sys.argv) > 2 and sys.argv[2] == "nograph":
    assert w._lib.dhg_set_option(w._ctx, b"graph", 0) == 0
g = torch.Generator().manual_seed(7)
B, T, L = 48, 392, 24
text = torch.randint(2, 73, (B, L), generator=g); text[:, -1] = 1; text[::3, 17:] = 0
style = torch.randn(B, 14, 1280, generator=g)
x0 = torch.randn(B, T, 2, generator=g)
noise = torch.randn(60, B, T, 2, generator=g)
out = w.sample(text, style, T=T, x0=x0, noise=noise)
np.save(sys.argv[1], out.float().cpu().numpy())
"""


def test_chain_is_bit_identical_with_and_without_text_overlap(tmp_path):
    """The text sides of 4 consecutive steps run at once on their own streams (engine.cu run_chain); one step at a
    time (text_sets=1), three at a time on real streams without a CUDA graph, and the chain with every kernel walking its rows
    first-to-last under the built-in tile rule (serpentine=0,autotune=0) must give the same bits."""
    import numpy as np

    code = _CHAIN.format(root=ROOT, pkg=os.path.join(ROOT, "diffusion-handwriting-generation.pytorch_b200"))
    outs = {}
    for name, opts, graph in [("overlap", "", "graph"), ("serial", "text_sets=1", "graph"), ("overlap_streams", "text_sets=3", "nograph"),
                              ("one_direction", "serpentine=0,autotune=0", "graph"), ("no_tail_fusion", "tail_fusion=0", "graph"), ("tail_fusion_1", "tail_fusion=1", "graph"), ("tail_fusion_2", "tail_fusion=2", "graph"),
                              ("no_head_fusion", "head_fusion=0", "graph"), ("no_skip_fusion", "skip_fusion=0", "graph"),
                              ("skip_fusion_enc_only", "skip_fusion=7", "graph"), ("skip_fusion_one_direction", "serpentine=0,autotune=0,skip_fusion=31", "graph"),
                              ("direct_store_few_slots", "direct_store=1,attn_max_slots=2,max_stages_a=3", "graph")]:
        f = str(tmp_path / f"{name}.npy")
        r = subprocess.run([sys.executable, "-c", code, f, graph], env=dict(os.environ, DHG_OPTS=opts), capture_output=True, text=True,
                           cwd=ROOT, timeout=900)
        assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
        outs[name] = np.load(f)
    assert np.isfinite(outs["overlap"]).all()
    assert np.array_equal(outs["overlap"], outs["serial"])
    assert np.array_equal(outs["overlap"], outs["overlap_streams"])
    assert np.array_equal(outs["overlap"], outs["one_direction"])   # tile order and tile configuration only move time
    assert np.array_equal(outs["overlap"], outs["skip_fusion_one_direction"])   # also for the dual-operand GEMMs
    assert np.array_equal(outs["overlap"], outs["direct_store_few_slots"])      # store path, slot count, ring depth: time only
    # the fused tail (fc + FiLM + skip + heads on folded fp32 tables) skips one bf16 rounding of d1: close, not identical
    # (likewise the fused head: enc1.conv_skip from x in fp32 instead of from the bf16 input_dense rows)
    # (likewise skip fusion: conv_skip accumulated inside the block's last GEMM in fp32 instead of through a bf16 `skip` row,
    # and the FiLM scale folded into bf16 weights)
    for other in ("no_tail_fusion", "tail_fusion_1", "tail_fusion_2", "no_head_fusion", "no_skip_fusion", "skip_fusion_enc_only"):
        a, b = outs["overlap"].astype(np.float64), outs[other].astype(np.float64)
        assert np.linalg.norm(a[..., :2] - b[..., :2]) / np.linalg.norm(b[..., :2]) < 1e-2, other
        assert np.abs(a[..., 2] - b[..., 2]).mean() < 2e-2 and np.abs(a[..., 2] - b[..., 2]).max() < 0.25, other   # pen probabilities (two bf16 variants of a 60-step chain)
