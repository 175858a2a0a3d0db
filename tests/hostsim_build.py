"""Host build of the training-step kernels (csrc/train_step.cu with -DDHG_HOSTSIM): TEST INFRASTRUCTURE ONLY.

The training step's kernels are written as functors over a flat index; compiled with g++ the same bodies run in a loop,
so the tape (op order, strides of every contraction, every backward formula) can be checked against torch autograd in
the CPU test suite.  Nothing under the product package loads this library; libdhg_b200.so does not contain it.
"""
import ctypes
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
SRC = os.path.join(ROOT, "diffusion-handwriting-generation.pytorch_b200", "csrc", "train_step.cu")
OUT = os.path.join(HERE, "_hostsim", "libdhg_train_hostsim.so")


def build():
    hdr = os.path.join(ROOT, "include", "dhg_b200.h")
    if os.path.exists(OUT) and os.path.getmtime(OUT) > max(os.path.getmtime(SRC), os.path.getmtime(hdr)):
        return OUT
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    cmd = ["g++", "-O2", "-fopenmp", "-std=c++17", "-shared", "-fPIC", "-DDHG_HOSTSIM", "-x", "c++", SRC, "-o", OUT]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError("host build of train_step.cu failed:\n" + r.stdout)
    return OUT


def lib():
    from dhg_b200 import _abi

    h = ctypes.CDLL(build())
    for name, (res, args) in _abi.SIGNATURES.items():
        if name.startswith("dhg_trainer_"):
            f = getattr(h, name)
            f.restype = res
            f.argtypes = args
    return h
