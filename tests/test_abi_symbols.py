"""The C-ABI library builds here (nvcc cross-compiles without a GPU), loads, and exports every
symbol include/dhg_b200.h declares.  No compute calls: there is no GPU on the CPU tier."""
import ctypes
import os
import re

import pytest
import torch

from dhg_b200 import _abi

HEADER = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "include", "dhg_b200.h")


def declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(dhg_[a-z0-9_]+)\s*\(", src)))


def test_every_declared_symbol_is_exported_and_bound(built_lib):
    names = declared_functions()
    assert len(names) >= 18
    for n in names:
        assert hasattr(built_lib, n), f"{n} declared in the header but not exported"
    assert sorted(_abi.SIGNATURES) == names, "ctypes table and header disagree"
    assert built_lib.dhg_abi_version() == 1


def test_library_is_sm100a_only(built_lib):
    import subprocess

    out = subprocess.run(["cuobjdump", "-lelf", _abi.LIB_PATH], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs


@pytest.mark.skipif(torch.cuda.is_available(), reason="CPU-tier behaviour")
def test_fails_loudly_without_gpu(built_lib, state_dict):
    ctx = ctypes.c_void_p(0)
    cfg = _abi.DhgConfig(2, 128)
    rc = built_lib.dhg_create(0, ctypes.byref(cfg), ctypes.byref(ctx))
    assert rc != 0 and b"no CPU fallback" in built_lib.dhg_last_error()
    from dhg_b200 import DiffusionWriter

    with pytest.raises(_abi.DhgError, match="no CPU fallback"):
        DiffusionWriter(state_dict=state_dict, num_layers=2, channels=128)


def test_bad_config_is_rejected(built_lib):
    ctx = ctypes.c_void_p(0)
    cfg = _abi.DhgConfig(2, 64)   # the reference's 32-wide sigma embedding only fits channels=128
    assert built_lib.dhg_create(0, ctypes.byref(cfg), ctypes.byref(ctx)) != 0
    assert b"channels must be 128" in built_lib.dhg_last_error()
    assert built_lib.dhg_create(0, None, ctypes.byref(ctx)) != 0
