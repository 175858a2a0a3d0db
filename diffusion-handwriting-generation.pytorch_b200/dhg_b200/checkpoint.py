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
