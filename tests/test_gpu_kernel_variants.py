"""The same GEMM unit cases under every kernel-selection switch (generic epilogue instance, forced CTA pairs
(cta_group::2), no interleaving / streaming weights): each variant runs in its own process because the switches
are read when the library is loaded (DHG_OPTS)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("opts", ["specialize=0", "pair=2", "pair=0,interleave=0,w_resident=0", "pdl=0"])
def test_gemm_cases_under_switch(opts):
    env = dict(os.environ, DHG_OPTS=opts)
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.join(ROOT, "tests", "test_gpu_gemm.py"), "-x", "-q", "-m", "gpu",
                        "-p", "no:cacheprovider"], env=env, capture_output=True, text=True, cwd=ROOT, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
