"""`infer()` with the reference's signature and checkpoint discovery
(inference.py:19-98), running the chain on the B200 engine.

`source` is the reference's writer image (read_img(source, 96) -> StyleExtractor,
inference.py:67-70; here `dhg_b200.style`, CUDA).  The reference downloads the
pretrained MobileNetV2 weights at that point; offline they come from `style_weights`
(a torchvision mobilenet_v2 state_dict or its path) or DHG_MOBILENET_WEIGHTS.
`source` may also be a `.pt`/`.npy` file holding a precomputed [14,1280] (or
[1,14,1280]) style tensor.  The PNG is rasterised by `vis.save_strokes_png` because
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


def load_style(source, style_weights=None, device="cuda:0"):
    if isinstance(source, torch.Tensor):
        s = source
    elif str(source).endswith(".npy"):
        import numpy as np

        s = torch.from_numpy(np.load(source))
    elif str(source).endswith((".pt", ".pth")):
        s = torch.load(source, map_location="cpu")
    else:   # a writer image: inference.py:67-70
        from .style import StyleExtractor, read_img

        extractor = StyleExtractor(style_weights, device=device)
        writer_img = read_img(source, 96)[None, None, :]
        s = extractor(writer_img).cpu()
        extractor.close()
    s = s.float()
    return s[None] if s.dim() == 2 else s


def infer(prompt, source, config_path=None, checkpoint_path=None, experiment_path=None, output="result",
          diffusion_mode="new", *, dtype="fp32", device="cuda:0", seed=None, style_weights=None):
    config_path, checkpoint_path = resolve_experiment(config_path, checkpoint_path, experiment_path)
    writer = DiffusionWriter(config_path, checkpoint_path, dtype=dtype, device=device)
    style = load_style(source, style_weights, device)
    ids = Tokenizer().encode(prompt)
    text = torch.tensor([ids])
    strokes = writer.sample(text, style, T=stroke_length(len(ids)), diffusion_mode=diffusion_mode, seed=seed)
    strokes = strokes[0].cpu()
    if output:
        from .vis import save_strokes_png

        save_strokes_png(strokes.numpy(), f"./{output}.png")
    return strokes


def main(argv=None):
    """Command line of the reference: `python diffusion_handwriting_generation/inference.py --prompt=... --source=...
    --experiment_path=... --output=...` (inference.py:101-102 `fire.Fire(infer)`, driven by `make infer TEXT= SOURCE=
    EXP= OUTPUT=`, Makefile:14-21).  `fire` is not installed here, so the same flags are parsed with argparse; both
    `--flag value` and fire's `--flag=value` spellings work, and the positional order is infer()'s."""
    import argparse

    ap = argparse.ArgumentParser(prog="python -m dhg_b200.inference", description=infer.__doc__ or main.__doc__)
    ap.add_argument("prompt_pos", nargs="?", default=None, help="text to write (or --prompt)")
    ap.add_argument("source_pos", nargs="?", default=None, help="style source (or --source)")
    ap.add_argument("--prompt", "--text", dest="prompt", default=None)
    ap.add_argument("--source", default=None, help="handwriting image, or a .pt/.npy file with a [14,1280] style tensor")
    ap.add_argument("--config_path", "--config-path", dest="config_path", default=None)
    ap.add_argument("--checkpoint_path", "--checkpoint-path", dest="checkpoint_path", default=None)
    ap.add_argument("--experiment_path", "--experiment-path", "--exp", dest="experiment_path", default=None)
    ap.add_argument("--output", default="result")
    ap.add_argument("--diffusion_mode", "--diffusion-mode", dest="diffusion_mode", default="new", choices=["new", "standard"])
    ap.add_argument("--dtype", default="fp32", choices=["fp32", "bf16"])
    ap.add_argument("--device", default="cuda:0")
    ap.add_argument("--seed", type=int, default=None)
    ap.add_argument("--style_weights", "--style-weights", dest="style_weights", default=None,
                    help="torchvision mobilenet_v2 weights file for an image --source (default: $DHG_MOBILENET_WEIGHTS)")
    a = ap.parse_args(argv)
    prompt = a.prompt if a.prompt is not None else a.prompt_pos
    source = a.source if a.source is not None else a.source_pos
    if prompt is None or source is None:
        ap.error("prompt and source are required (infer(prompt, source, ...), inference.py:19-27)")
    strokes = infer(prompt, source, a.config_path, a.checkpoint_path, a.experiment_path, a.output, a.diffusion_mode,
                    dtype=a.dtype, device=a.device, seed=a.seed, style_weights=a.style_weights)
    print(f"wrote ./{a.output}.png ({strokes.shape[0]} stroke points)")
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
