"""`StyleExtractor`, `read_img`, `remove_whitespace`: the step right before the sampling path, with the
reference's names and argument meaning (text_style.py:11-59, utils/io.py:98-115, utils/preprocessing.py:47-62).

The network is torchvision's MobileNetV2 feature stack; its forward runs in hand-written fp32 CUDA kernels through the
C ABI (`dhg_style_*`, csrc/style_extractor.cu).  The reference downloads the pretrained ImageNet weights when the class
is constructed; there is no network here, so the weights are an argument: a torchvision `mobilenet_v2` state_dict (or
the path of one, e.g. torchvision's `mobilenet_v2-7ebf99e0.pth`), or the `DHG_MOBILENET_WEIGHTS` environment variable.
"""
import ctypes
import os

import numpy as np
import torch

from . import _abi


def remove_whitespace(img, thresh, remove_middle=False):
    """utils/preprocessing.py:47-62: drop the rows / columns without a pixel darker than `thresh` (outer ones only
    unless `remove_middle`).  Like the reference, the last dark row / column itself is cut too (`rows[0]:rows[-1]`)."""
    row_mins, col_mins = np.amin(img, axis=1), np.amin(img, axis=0)
    rows, cols = np.nonzero(row_mins < thresh)[0], np.nonzero(col_mins < thresh)[0]
    if remove_middle:
        return img[rows][:, cols]
    return img[rows[0]:rows[-1], cols[0]:cols[-1]]


def read_img(path, height):
    """utils/io.py:98-115: grey image, whitespace removed, resized to `height` rows (bicubic), uint8 [height, W]."""
    import cv2   # the reference's own dependency for this step

    img = cv2.imread(os.fspath(path), cv2.IMREAD_GRAYSCALE)
    if img is None:
        raise FileNotFoundError(f"cannot read image {path}")
    img = remove_whitespace(img, thresh=127)
    h, w = img.shape
    return cv2.resize(img, (height * w // h, height), interpolation=cv2.INTER_CUBIC)


def _load_weights(weights):
    if weights is None:
        weights = os.environ.get("DHG_MOBILENET_WEIGHTS")
        if not weights:
            raise ValueError(
                "StyleExtractor needs the MobileNetV2 weights the reference downloads (torchvision MobileNet_V2_Weights.DEFAULT): "
                "pass weights=<state_dict or path of mobilenet_v2-*.pth> or set DHG_MOBILENET_WEIGHTS; there is no network here")
    if isinstance(weights, (str, os.PathLike)):
        weights = torch.load(os.fspath(weights), map_location="cpu", weights_only=True)
    if hasattr(weights, "state_dict"):
        weights = weights.state_dict()
    return weights


class StyleExtractor:
    """Extracts style features [B, 14, 1280] from grey handwriting images (text_style.py:11-59)."""

    def __init__(self, weights=None, device="cuda:0"):
        sd = _load_weights(weights)
        self.device = torch.device(device)
        if self.device.type != "cuda" or not torch.cuda.is_available():
            raise _abi.DhgError("StyleExtractor needs a CUDA device: there is no CPU fallback")
        index = self.device.index if self.device.index is not None else torch.cuda.current_device()
        self.device = torch.device("cuda", index)
        self._lib = _abi.lib()
        self._h = ctypes.c_void_p(0)
        self._check(self._lib.dhg_style_create(index, ctypes.byref(self._h)))
        for name, t in sd.items():
            if not name.startswith("features.") or name.endswith("num_batches_tracked"):
                continue
            t = t.detach().to("cpu", torch.float32).contiguous()
            dims = (ctypes.c_int64 * max(t.dim(), 1))(*t.shape)
            self._check(self._lib.dhg_style_load_weight(self._h, name.encode(), ctypes.c_void_p(t.data_ptr()), dims, t.dim()))
        self._check(self._lib.dhg_style_finalize(self._h))

    def _check(self, rc):
        if rc != 0:
            raise _abi.DhgError(self._lib.dhg_style_last_error().decode("utf-8", "replace"))

    @torch.no_grad()
    def __call__(self, img_batch):
        """img_batch: [B, 1, H, W] (or [B, H, W]) grey levels 0..255, numpy or tensor -> [B, 14, 1280] on the device."""
        x = torch.as_tensor(np.asarray(img_batch) if not isinstance(img_batch, torch.Tensor) else img_batch)
        x = x.to("cpu", torch.float32)
        if x.dim() == 4:
            if x.shape[1] != 1:
                raise ValueError("img_batch must be [B,1,H,W] grey images")
            x = x[:, 0]
        if x.dim() != 3:
            raise ValueError("img_batch must be [B,1,H,W] or [B,H,W]")
        x = x.contiguous()
        B, H, W = x.shape
        out = torch.empty(B, 14, 1280, device=self.device, dtype=torch.float32)
        with torch.cuda.device(self.device):
            stream = ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
            self._check(self._lib.dhg_style_extract(self._h, ctypes.c_void_p(x.data_ptr()), B, H, W, ctypes.c_void_p(out.data_ptr()), stream))
            torch.cuda.current_stream(self.device).synchronize()   # the host image is borrowed until the copy has happened
        return out

    forward = __call__

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            self._lib.dhg_style_destroy(self._h)
            self._h = ctypes.c_void_p(0)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
