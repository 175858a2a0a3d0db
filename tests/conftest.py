import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "diffusion-handwriting-generation.pytorch_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


@pytest.fixture(scope="session")
def state_dict():
    from oracle.dhg_oracle import init_state_dict

    return init_state_dict(0)


@pytest.fixture(scope="session")
def golden():
    import numpy as np

    def load(name):
        return np.load(os.path.join(GOLDEN, name + ".npz"))

    return load


@pytest.fixture(scope="session")
def built_lib():
    """The C-ABI library, built in-tree if needed (nvcc cross-compiles without a GPU)."""
    from dhg_b200 import _abi
    from dhg_b200 import build as _build

    _build.build()
    return _abi.lib()
