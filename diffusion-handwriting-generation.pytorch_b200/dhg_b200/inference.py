"""`infer()` with the reference's signature and checkpoint discovery
(inference.py:19-98), running the chain on the B200 engine.

Differences forced by the environment, not by design: the reference derives the
style vector from a handwriting image with a pretrained MobileNetV2
(text_style.py:11-59) whose weights cannot be fetched offline, so `source` may
also be a `.pt`/`.npy` file holding a precomputed [14,1280] (or [1,14,1280])
style tensor; and the PNG is rasterised by `vis.save_strokes_png` because
matplotlib is not installed.
"""
from pathlib import Path

import torch

from .tokenizer import Tokenizer, stroke_length
from .writer import DiffusionWriter


def resolve_experiment(config_path=None, checkpoint_path=None, experiment_path=None):
    """Checkpoint search order of inference.py:28-58: model_final.pth, model_last.pth,
    then the highest-numbered checkpoint_<int>.pth."""
    if experiment_path:
        exp = Path(experiment_path)
        if not config_path:
            config_path = str(exp / "config.yml")
        if not checkpoint_path:
            ckpt = exp / "model_final.pth"
            if not ckpt.exists():
                ckpt = exp / "model_last.pth"
            if not ckpt.exists():
                numbered = []
                for p in exp.glob("checkpoint_*.pth"):
                    try:
                        numbered.append((int(p.stem.split("_")[1]), p))
                    except ValueError:
                        continue
                if numbered:
                    ckpt = max(numbered, key=lambda sp: sp[0])[1]
            if ckpt.exists():
                checkpoint_path = str(ckpt)
    if not config_path or not checkpoint_path:
        raise ValueError(
            "Both config_path and checkpoint_path must be provided, either directly or via experiment_path."
        )
    return config_path, checkpoint_path


def load_style(source):
    if isinstance(source, torch.Tensor):
        s = source
    elif str(source).endswith(".npy"):
        import numpy as np

        s = torch.from_numpy(np.load(source))
    elif str(source).endswith((".pt", ".pth")):
        s = torch.load(source, map_location="cpu")
    else:
        raise NotImplementedError(
            "style extraction from an image needs the pretrained MobileNetV2 of the reference's "
            "StyleExtractor, which is outside this path; pass a precomputed [14,1280] style tensor (.pt/.npy)"
        )
    s = s.float()
    return s[None] if s.dim() == 2 else s


def infer(prompt, source, config_path=None, checkpoint_path=None, experiment_path=None, output="result",
          diffusion_mode="new", *, dtype="fp32", device="cuda:0", seed=None):
    config_path, checkpoint_path = resolve_experiment(config_path, checkpoint_path, experiment_path)
    writer = DiffusionWriter(config_path, checkpoint_path, dtype=dtype, device=device)
    style = load_style(source)
    ids = Tokenizer().encode(prompt)
    text = torch.tensor([ids])
    strokes = writer.sample(text, style, T=stroke_length(len(ids)), diffusion_mode=diffusion_mode, seed=seed)
    strokes = strokes[0].cpu()
    if output:
        from .vis import save_strokes_png

        save_strokes_png(strokes.numpy(), f"./{output}.png")
    return strokes
