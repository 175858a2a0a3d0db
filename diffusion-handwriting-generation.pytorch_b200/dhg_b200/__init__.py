"""dhg_b200 -- B200-native reverse-diffusion sampling for the handwriting model
of sleep3r/Diffusion-Handwriting-Generation.pytorch.

Host-side mirror of the reference's call surface for this path:
`infer`, `load_model`, `DiffusionModel.forward`, `Tokenizer`, `get_beta_set`,
plus the `DiffusionWriter` facade.  All compute goes through the C ABI in
include/dhg_b200.h (lib/libdhg_b200.so, hand-written sm_100a kernels); there is
no CPU fallback.
"""
from .checkpoint import DiffusionModel, load_model, save_checkpoint, save_model_final
from .config import DLConfig
from .diffusion import NUM_STEPS, get_alpha_bar, get_beta_set
from .inference import infer, resolve_experiment
from .style import StyleExtractor, read_img, remove_whitespace
from .tokenizer import Tokenizer, stroke_length
from .train import FlatAdam, InvSqrtSchedule, exchange_gradients, loss_fn, perturb
from .writer import DiffusionWriter

__all__ = [
    "DiffusionWriter", "DiffusionModel", "load_model", "infer", "resolve_experiment", "Tokenizer",
    "stroke_length", "DLConfig", "save_checkpoint", "save_model_final", "StyleExtractor", "read_img", "remove_whitespace", "get_beta_set", "get_alpha_bar", "NUM_STEPS",
    "FlatAdam", "InvSqrtSchedule", "exchange_gradients", "loss_fn", "perturb",
]
