"""Recipe for `baseline/_ref/`: the UNMODIFIED reference package, importable on the GPU box
(where /root/reference does not exist), for bench.py's `--impl reference` arm and its same-box
GPU-eager baseline.

    python baseline/install_ref.py            # run here, in the build container

1. The contract's install: `pip install --no-index --no-build-isolation --find-links /opt/wheelhouse
   --target baseline/_ref <copy of /root/reference>`.  Outcome in this image: it fails, the reference's
   build backend (`hatchling`, pyproject.toml:27-29) is in neither the venv nor the wheelhouse.
2. So the recipe does what that wheel would have done: hatchling's only instruction is
   `packages = ["diffusion_handwriting_generation"]` (pyproject.toml:31-32), a pure-Python wheel, i.e. a
   verbatim copy of that one package directory into the target.  Nothing is edited.
`baseline/_ref/` is git-ignored (the reference's sources never enter this repo's history) but not
gpurun-ignored, so it travels to the GPU box.  Only `bench.py` (reference arm, eager GPU baseline) imports it.
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
TARGET = os.path.join(HERE, "_ref")
REFERENCE = "/root/reference"
PACKAGE = "diffusion_handwriting_generation"


def install(verbose=True):
    if not os.path.isdir(os.path.join(REFERENCE, PACKAGE)):
        return os.path.isdir(os.path.join(TARGET, PACKAGE))   # GPU box: use what travelled with the tree
    note = []
    tmp = "/tmp/dhg_ref_copy"
    shutil.rmtree(tmp, ignore_errors=True)
    shutil.copytree(REFERENCE, tmp, ignore=shutil.ignore_patterns(".git", "data"))
    cmd = [sys.executable, "-m", "pip", "install", "--no-index", "--no-build-isolation", "--no-deps", "--find-links",
           "/opt/wheelhouse", "--target", TARGET, tmp]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode == 0 and os.path.isdir(os.path.join(TARGET, PACKAGE)):
        note.append("pip install --target succeeded")
    else:
        last = (r.stdout or "").strip().splitlines()[-1:] or ["?"]
        note.append(f"pip install failed ({last[0]}); copied the package directory verbatim instead")
        shutil.rmtree(os.path.join(TARGET, PACKAGE), ignore_errors=True)
        os.makedirs(TARGET, exist_ok=True)
        shutil.copytree(os.path.join(REFERENCE, PACKAGE), os.path.join(TARGET, PACKAGE),
                        ignore=shutil.ignore_patterns("__pycache__"))
    shutil.rmtree(tmp, ignore_errors=True)
    with open(os.path.join(TARGET, "_INSTALL_NOTE.txt"), "w") as f:
        f.write("\n".join(note) + "\n")
    if verbose:
        print("baseline/_ref:", "; ".join(note))
    return True


def import_reference():
    """-> (DiffusionModel, new_diffusion_step, get_beta_set) of the unmodified reference, or None."""
    if not os.path.isdir(os.path.join(TARGET, PACKAGE)):
        return None
    if TARGET not in sys.path:
        sys.path.insert(0, TARGET)
    from diffusion_handwriting_generation.model import DiffusionModel
    from diffusion_handwriting_generation.utils.nn import get_beta_set, new_diffusion_step

    return DiffusionModel, new_diffusion_step, get_beta_set


if __name__ == "__main__":
    ok = install()
    print("importable:", import_reference() is not None and ok)
