"""`DiffusionWriter`: the host-side facade over the sm_100a engine.

It performs exactly what the reference's `infer()` does between loading the
model and plotting (inference.py:60-96), for any batch size, through the C ABI
in include/dhg_b200.h.  PyTorch is used for device memory and streams only.
"""
import ctypes
import weakref
import re

import torch

from . import _abi
from .config import DLConfig
from .diffusion import NUM_STEPS, get_alpha_bar, get_beta_set
from .tokenizer import Tokenizer, stroke_length

_PREC = {"fp32": 0, "float32": 0, torch.float32: 0, "bf16": 1, "bfloat16": 1, torch.bfloat16: 1}
_MODE = {"new": 0, "standard": 1}
STYLE_WIDTH = 1280


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p(0)


def read_state_dict(checkpoint_path):
    """checkpoint.py:117-129: accept a raw state_dict or {"state_dict": ...}, strip a
    leading `module.` from every key."""
    ckpt = torch.load(checkpoint_path, map_location="cpu", weights_only=False)
    if not isinstance(ckpt, dict):
        raise RuntimeError(f"No state_dict found in checkpoint file {checkpoint_path}")
    sd = ckpt["state_dict"] if "state_dict" in ckpt else ckpt
    return {re.sub(r"^module\.", "", k): v for k, v in sd.items()}


class DiffusionWriter:
    """Reverse-diffusion handwriting sampler on one B200.

    Args:
        config_path: experiment `config.yml` (reads training_args.att_layers_num /
            .channels like checkpoint.py:280-286).  Optional if `num_layers` /
            `channels` are given.
        checkpoint_path: `model_final.pth`-style file, or pass `state_dict`.
        device: CUDA device (there is no CPU path).
        dtype: "bf16" (bf16 storage, tcgen05 tensor cores, fp32 accumulate) or "fp32" (the reference's precision:
            parity 1e-3 after the chain.  Runs on the tensor cores too, on split bf16 hi/lo storage with three-term
            products; `gemm=0` selects plain fp32 storage with CUDA-core FMA GEMMs instead).
        chunk: samples per captured chain; larger batches run as several chunks.
    """

    def __init__(self, config_path=None, checkpoint_path=None, *, state_dict=None, num_layers=None,
                 channels=None, device="cuda:0", dtype="bf16", chunk=1024, cfg_options=None,
                 gemm=None, graph=None):
        if dtype not in _PREC:
            raise ValueError(f"dtype must be 'fp32' or 'bf16', got {dtype!r}")
        self.precision = _PREC[dtype]
        self.dtype = "fp32" if self.precision == 0 else "bf16"
        if config_path is not None:
            cfg = DLConfig.load(config_path)
            cfg.update(cfg_options)
            num_layers = cfg.training_args.att_layers_num if num_layers is None else num_layers
            channels = cfg.training_args.channels if channels is None else channels
        if num_layers is None or channels is None:
            raise ValueError("need config_path, or num_layers and channels")
        if state_dict is None:
            if checkpoint_path is None:
                raise ValueError("need checkpoint_path or state_dict")
            state_dict = read_state_dict(checkpoint_path)
        else:
            state_dict = {re.sub(r"^module\.", "", k): v for k, v in state_dict.items()}

        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise _abi.DhgError("DiffusionWriter needs a CUDA device: there is no CPU fallback")
        if not torch.cuda.is_available():
            raise _abi.DhgError("CUDA is not available: DiffusionWriter has no CPU fallback")
        self._lib = _abi.lib()
        self.num_layers, self.channels = int(num_layers), int(channels)
        self.chunk = int(chunk)
        self._ctx = ctypes.c_void_p(0)
        cfg_c = _abi.DhgConfig(self.num_layers, self.channels)
        index = self.device.index if self.device.index is not None else torch.cuda.current_device()
        self.device = torch.device("cuda", index)
        _abi.check(self._lib.dhg_create(index, ctypes.byref(cfg_c), ctypes.byref(self._ctx)))
        if gemm is not None:
            _abi.check(self._lib.dhg_set_option(self._ctx, b"gemm", int(gemm)))
        if graph is not None:
            _abi.check(self._lib.dhg_set_option(self._ctx, b"graph", int(graph)))
        self._load(state_dict)
        self._plan_key = None
        self.tokenizer = Tokenizer()

    # -- weights -----------------------------------------------------------
    def expected_keys(self):
        n = self._lib.dhg_num_weights(self._ctx)
        out = {}
        for i in range(n):
            name = self._lib.dhg_weight_name(self._ctx, i).decode()
            nd = self._lib.dhg_weight_ndim(self._ctx, i)
            out[name] = tuple(self._lib.dhg_weight_dim(self._ctx, i, d) for d in range(nd))
        return out

    def _load(self, state_dict):
        expected = self.expected_keys()
        unexpected = [k for k in state_dict if k not in expected]
        missing = [k for k in expected if k not in state_dict]
        if unexpected or missing:  # strict=True semantics of checkpoint.py:83-87
            msg = ["The model and loaded state dict do not match exactly"]
            if unexpected:
                msg.append("unexpected key in source state_dict: " + ", ".join(unexpected[:8]))
            if missing:
                msg.append("missing keys in source state_dict: " + ", ".join(missing[:8]))
            raise RuntimeError("\n".join(msg))
        for name, shape in expected.items():
            t = state_dict[name].detach().to("cpu", torch.float32).contiguous()
            if tuple(t.shape) != shape:
                raise RuntimeError(f"size mismatch for {name}: expected {shape}, got {tuple(t.shape)}")
            dims = (ctypes.c_int64 * len(shape))(*shape)
            _abi.check(self._lib.dhg_load_weight(self._ctx, name.encode(), _ptr(t), dims, len(shape)))
        beta = get_beta_set().to(torch.float32).contiguous()
        abar = get_alpha_bar(beta).to(torch.float32).contiguous()
        _abi.check(self._lib.dhg_set_schedule(self._ctx, _ptr(beta), _ptr(abar)))
        _abi.check(self._lib.dhg_finalize(self._ctx))
        self.beta, self.alpha_bar = beta, abar

    # -- helpers -------------------------------------------------------------
    def _plan(self, B, T, L, S):
        key = (B, T, L, S, self.precision)
        if key != self._plan_key:
            _abi.check(self._lib.dhg_plan(self._ctx, B, T, L, S, self.precision))
            self._plan_key = key

    def _dev(self, t, dtype):
        t = torch.as_tensor(t)
        return t.to(device=self.device, dtype=dtype).contiguous()

    def _stream(self):
        return ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def encode(self, prompts):
        """list[str] -> right-padded int64 [B, L] (pad id 0) using the reference tokenizer."""
        ids = [self.tokenizer.encode(p) for p in prompts]
        L = max(len(i) for i in ids)
        out = torch.zeros(len(ids), L, dtype=torch.int64)
        for r, i in enumerate(ids):
            out[r, : len(i)] = torch.tensor(i)
        return out

    def _check_tokens(self, text, original=None):
        """Out-of-range ids raise IndexError like the reference's nn.Embedding.  Checking a CUDA tensor reads the
        device (a synchronisation), so the very same tensor OBJECT, unmodified since it last passed (torch's version
        counter), is not checked again.  `original`: the caller's tensor before any device copy."""
        obj = original if isinstance(original, torch.Tensor) else text
        ok = getattr(self, "_tokens_ok", None)
        if ok is not None and ok[0]() is obj and ok[1] == obj._version:
            return
        if text.numel() and (int(text.min()) < 0 or int(text.max()) >= 73):
            raise IndexError("text token id out of range [0, 73)")
        self._tokens_ok = (weakref.ref(obj), obj._version)

    @staticmethod
    def _check_mode(diffusion_mode):
        if diffusion_mode not in _MODE:
            raise ValueError(f"diffusion_mode must be 'new' or 'standard', got {diffusion_mode!r}")
        return _MODE[diffusion_mode]

    @staticmethod
    def _check_chain_shapes(text, style, x0, noise):
        """Shape contract of the chain entry points (sample / sample_host): every pointer handed to the C ABI is
        sized from these, so nothing is passed on before they hold."""
        if text.dim() != 2:
            raise ValueError("text must be [B,L]")
        B, L = text.shape
        if style.dim() != 3 or style.shape[0] != B or style.shape[2] != STYLE_WIDTH or style.shape[1] < 1:
            raise ValueError("style_vector must be [B,S,1280]")
        if x0.dim() != 3 or tuple(x0.shape[::2]) != (B, 2):
            raise ValueError("x0 must be [B,T,2] and noise [60,B,T,2]")
        T = x0.shape[1]
        if T % 8 or T <= 0:
            raise ValueError("T must be a positive multiple of 8")
        if tuple(noise.shape) != (NUM_STEPS, B, T, 2):
            raise ValueError("x0 must be [B,T,2] and noise [60,B,T,2]")
        return B, T, L, style.shape[1]

    def check_errors(self):
        """Synchronise the current stream and raise IndexError if a stream-ordered call met a token id outside
        [0, 73) (the reference's nn.Embedding raises there; the kernel can only flag it)."""
        with torch.cuda.device(self.device):
            rc = self._lib.dhg_check_errors(self._ctx, self._stream())
        if rc != 0:
            msg = self._lib.dhg_last_error().decode("utf-8", "replace")
            if "token id" in msg:
                raise IndexError(msg)
            raise _abi.DhgError(msg)

    # -- DiffusionModel.forward (model.py:121-182) -----------------------------
    @torch.no_grad()
    def denoise(self, strokes, text, sigma, style_vector):
        """-> (eps [B,T,2], pen_lifts [B,T], None).  `sigma` may be [B,1,1], [B,1] or [B]."""
        strokes = self._dev(strokes, torch.float32)
        B, T, two = strokes.shape
        if two != 2:
            raise ValueError("strokes must be [B,T,2]")
        text_in = text
        text = self._dev(text, torch.int64)
        self._check_tokens(text, text_in)
        style = self._dev(style_vector, torch.float32)
        if style.dim() != 3 or style.shape[0] != B or style.shape[2] != STYLE_WIDTH:
            raise ValueError("style_vector must be [B,S,1280]")
        sigma = self._dev(sigma, torch.float32).reshape(-1)
        if sigma.numel() != B or text.shape[0] != B:
            raise ValueError("batch size mismatch between strokes, text, sigma")
        self._plan(B, T, text.shape[1], style.shape[1])
        eps = torch.empty(B, T, 2, device=self.device, dtype=torch.float32)
        pen = torch.empty(B, T, device=self.device, dtype=torch.float32)
        with torch.cuda.device(self.device):
            _abi.check(self._lib.dhg_denoise(self._ctx, _ptr(strokes), _ptr(text), _ptr(sigma), _ptr(style),
                                             _ptr(eps), _ptr(pen), self._stream()))
        return eps, pen, None

    __call__ = denoise

    # -- the 60-step chain (inference.py:81-96) --------------------------------
    @torch.no_grad()
    def sample(self, text, style_vector, *, T=None, x0=None, noise=None, seed=None, diffusion_mode="new"):
        """Sample strokes [B,T,3] = (dx, dy, pen-lift probability).

        text: list[str] | int tensor [B,L] (0 = padding).  style_vector: [B,S,1280].
        T: stroke points; default from the prompt length like inference.py:77-78.
        x0 [B,T,2] / noise [60,B,T,2]: injected initial state / per-step draws
        (noise[i] is consumed at loop index i); by default drawn with torch's CUDA
        generator (seeded with `seed` if given).
        """
        mode = self._check_mode(diffusion_mode)
        if isinstance(text, (list, tuple)) and text and isinstance(text[0], str):
            text = self.encode(text)
        text_in = text
        text = self._dev(text, torch.int64)
        if text.dim() != 2:
            raise ValueError("text must be [B,L]")
        self._check_tokens(text, text_in)
        B, L = text.shape
        style = self._dev(style_vector, torch.float32)
        if style.dim() != 3 or style.shape[0] != B or style.shape[2] != STYLE_WIDTH:
            raise ValueError("style_vector must be [B,S,1280]")
        if T is None:
            T = x0.shape[1] if x0 is not None else stroke_length(L)
        if T % 8 or T <= 0:
            raise ValueError("T must be a positive multiple of 8")
        gen = None
        if seed is not None:
            gen = torch.Generator(device=self.device)
            gen.manual_seed(int(seed))
        if x0 is None:
            x0 = torch.randn(B, T, 2, device=self.device, generator=gen)
        if noise is None:
            noise = torch.randn(NUM_STEPS, B, T, 2, device=self.device, generator=gen)
        x0 = self._dev(x0, torch.float32)
        noise = self._dev(noise, torch.float32)
        if tuple(x0.shape) != (B, T, 2):
            raise ValueError("x0 must be [B,T,2] and noise [60,B,T,2]")
        self._check_chain_shapes(text, style, x0, noise)
        self._plan(min(B, self.chunk), T, L, style.shape[1])
        out = torch.empty(B, T, 3, device=self.device, dtype=torch.float32)
        with torch.cuda.device(self.device):
            _abi.check(self._lib.dhg_sample(self._ctx, B, _ptr(x0), _ptr(noise), 0, _ptr(text), _ptr(style),
                                            mode, _ptr(out), self._stream()))
        return out

    @torch.no_grad()
    def sample_host(self, text, style_vector, x0, noise, diffusion_mode="new", out=None):
        """Same chain through the HOST-buffer entry point (`dhg_sample_host`): CPU tensors
        in, CPU tensor out, all copies inside the call.  The device->host copy lands in a pinned
        staging buffer owned by the writer (allocating pinned memory per call costs tens of
        milliseconds every time the caching host allocator runs dry); the result is returned as a
        fresh tensor, or written into `out` ([B,T,3] fp32 CPU; pinned: no staging copy at all)."""
        mode = self._check_mode(diffusion_mode)
        if isinstance(text, (list, tuple)) and text and isinstance(text[0], str):
            text = self.encode(text)
        text = torch.as_tensor(text).to("cpu", torch.int64).contiguous()
        style = torch.as_tensor(style_vector).to("cpu", torch.float32).contiguous()
        x0 = torch.as_tensor(x0).to("cpu", torch.float32).contiguous()
        noise = torch.as_tensor(noise).to("cpu", torch.float32).contiguous()
        B, T, L, S = self._check_chain_shapes(text, style, x0, noise)
        # host ids: always checked (no device synchronisation involved); the library checks them once more
        if text.numel() and (int(text.min()) < 0 or int(text.max()) >= 73):
            raise IndexError("text token id out of range [0, 73)")
        self._plan(min(B, self.chunk), T, L, S)
        direct = out is not None and out.is_pinned() and out.is_contiguous()
        if out is not None and (tuple(out.shape) != (B, T, 3) or out.dtype != torch.float32 or out.device.type != "cpu"):
            raise ValueError("out must be a [B,T,3] fp32 CPU tensor")
        if direct:
            stage = out
        else:
            stage = getattr(self, "_host_stage", None)
            if stage is None or tuple(stage.shape) != (B, T, 3):
                stage = self._host_stage = torch.empty(B, T, 3, dtype=torch.float32, pin_memory=True)
        with torch.cuda.device(self.device):
            _abi.check(self._lib.dhg_sample_host(self._ctx, B, _ptr(x0), _ptr(noise), 0, _ptr(text), _ptr(style),
                                                 mode, _ptr(stage)))
        if direct:
            return out
        if out is not None:
            out.copy_(stage)
            return out
        return stage.clone()

    # -- new_diffusion_step / standard_diffusion_step (utils/nn.py:64-112) -----
    @torch.no_grad()
    def posterior_step(self, step, xt, eps, noise=None, diffusion_mode="new", out=None):
        xt = self._dev(xt, torch.float32)
        eps = self._dev(eps, torch.float32)
        noise = None if noise is None else self._dev(noise, torch.float32)
        out = torch.empty_like(xt) if out is None else out
        with torch.cuda.device(self.device):
            _abi.check(self._lib.dhg_posterior_step(self._ctx, int(step), _MODE[diffusion_mode], _ptr(xt), _ptr(eps),
                                                    _ptr(noise), _ptr(out), xt.numel(), self._stream()))
        return out

    def debug_read(self, name):
        """Test hook: fp32 copy [B, positions, C] of a named intermediate of the last forward."""
        n = int(self._lib.dhg_debug_read(self._ctx, name.encode(), ctypes.c_void_p(0), 0))
        if n < 0:
            _abi.check(1)
        buf = torch.empty(n, dtype=torch.float32)
        if int(self._lib.dhg_debug_read(self._ctx, name.encode(), _ptr(buf), n)) < 0:
            _abi.check(1)
        return buf.reshape(self._plan_key[0], -1)

    # -- introspection -----------------------------------------------------------
    @property
    def last_launch_count(self):
        return int(self._lib.dhg_last_launch_count(self._ctx))

    @property
    def plan_bytes(self):
        return int(self._lib.dhg_plan_bytes(self._ctx))

    def close(self):
        if getattr(self, "_ctx", None) and self._ctx.value:
            self._lib.dhg_destroy(self._ctx)
            self._ctx = ctypes.c_void_p(0)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
