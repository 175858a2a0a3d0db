"""In-tree build of the sm_100a C-ABI library (`lib/libdhg_b200.so`) with nvcc.

    python -m dhg_b200.build            # or __graft_entry__.build()

nvcc cross-compiles without a GPU.  The .so is git-ignored but travels with the
tree to the GPU box.
"""
import os
import shutil
import subprocess
import sys

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
PROJ_DIR = os.path.dirname(PKG_DIR)
CSRC = os.path.join(PROJ_DIR, "csrc")
LIB_DIR = os.path.join(PROJ_DIR, "lib")
LIB_PATH = os.environ.get("DHG_LIB_PATH") or os.path.join(LIB_DIR, "libdhg_b200.so")   # DHG_LIB_PATH: A/B builds
SOURCES = ["engine.cu", "kernels_simt.cu", "gemm_tc.cu", "attention_tc.cu", "style_extractor.cu", "train_update.cu", "train_step.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "--shared", "-Xcompiler", "-fPIC", "--use_fast_math=false",
]


def _nvcc():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; the dhg_b200 library cannot be built")


def _stale():
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)]
    deps.append(os.path.join(os.path.dirname(PROJ_DIR), "include", "dhg_b200.h"))
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def build(force=False, verbose=False):
    if not force and not _stale():
        return LIB_PATH
    os.makedirs(LIB_DIR, exist_ok=True)
    srcs = [os.path.join(CSRC, s) for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]
    flags = [f for f in NVCC_FLAGS if not f.startswith("--use_fast_math")] + os.environ.get("DHG_NVCC_FLAGS", "").split()
    objs = []
    procs = []
    for s in srcs:  # compile translation units in parallel
        o = os.path.join(LIB_DIR, os.path.basename(s) + ".o")
        objs.append(o)
        cmd = [_nvcc()] + [f for f in flags if f != "--shared"] + ["-c", s, "-o", o]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        procs.append((cmd, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for cmd, p in procs:
        out, _ = p.communicate()
        if verbose and out:
            print(out)
        if p.returncode != 0:
            raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + (out or ""))
    link = [_nvcc(), "--shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB_PATH] + objs
    r = subprocess.run(link, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n" + r.stdout)
    return LIB_PATH


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
