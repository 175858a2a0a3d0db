"""`load_model` with the reference's signature (checkpoint.py:256-297): returns
`(model, device)` where `model(strokes, text, sigma, style_vector)` behaves like
`DiffusionModel.forward` in eval mode, computed by the sm_100a engine."""
from .writer import DiffusionWriter, read_state_dict  # noqa: F401


class DiffusionModel:
    """Callable stand-in for the reference nn.Module on the sampling path
    (model.py:61-182).  Inference only: `.eval()` / `.to()` are accepted no-ops."""

    def __init__(self, writer):
        self.writer = writer

    def forward(self, strokes, text, sigma, style_vector):
        return self.writer.denoise(strokes, text, sigma, style_vector)

    __call__ = forward

    def eval(self):
        return self

    def to(self, *_a, **_k):
        return self


def load_model(config_path, checkpoint_path, cfg_options=None, *, dtype="fp32", device="cuda:0"):
    if checkpoint_path is None:
        raise ValueError("checkpoint_path is required: the engine has no random-init mode")
    writer = DiffusionWriter(config_path, checkpoint_path, cfg_options=cfg_options, dtype=dtype, device=device)
    return DiffusionModel(writer), "cuda"


def save_checkpoint(model, filename, meta=None, optimizer=None):
    """The reference's checkpoint file (checkpoint.py:225-253, written every `save_freq` steps by train.py:123-126):
    `{"meta": meta, "state_dict": <CPU tensors>, "optimizer": <optimizer.state_dict()>}` -- `optimizer` only when given
    (an optimizer, or a dict of named optimizers).  `model`: an nn.Module (a DataParallel-style `.module` wrapper is
    unwrapped) or a plain state_dict.  `read_state_dict` / `load_model` read it back."""
    import torch

    if hasattr(model, "module"):
        model = model.module
    sd = model.state_dict() if hasattr(model, "state_dict") else model
    checkpoint = {"meta": meta, "state_dict": type(sd)((k, v.detach().cpu()) for k, v in sd.items())}
    if optimizer is not None:
        if isinstance(optimizer, dict):
            checkpoint["optimizer"] = {name: opt.state_dict() for name, opt in optimizer.items()}
        else:
            checkpoint["optimizer"] = optimizer.state_dict()
    torch.save(checkpoint, filename)


def save_model_final(model, filename):
    """`model_final.pth` / `model_last.pth` as train.py:130-137 writes them: the raw state_dict, nothing around it."""
    import torch

    sd = model.state_dict() if hasattr(model, "state_dict") else model
    torch.save(type(sd)((k, v.detach().cpu()) for k, v in sd.items()), filename)
