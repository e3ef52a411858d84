"""Device-side image pre-processing: the reference datasets' per-image transform (dataset/__init__.py:150-159)

    transforms.Compose([transforms.Resize((S, S), Image.BICUBIC), transforms.ToTensor(),
                        transforms.Normalize(mean=CLIP_MEAN, std=CLIP_STD)])

for a batch of decoded RGB images, executed by libaaclip_b200.so (``aaclip_preprocess_u8``) and bit-exact with
PIL + torchvision. The host computes only the tap tables (``aaclip_resample_coeffs_host``, once per image size).
There is no CPU fallback: the images must be (or are copied) on a CUDA device.
"""
from __future__ import annotations

import ctypes
from typing import Dict, Sequence, Tuple

import numpy as np
import torch

from . import _lib

CLIP_MEAN = (0.48145466, 0.4578275, 0.40821073)     # dataset/__init__.py:154
CLIP_STD = (0.26862954, 0.26130258, 0.27577711)     # dataset/__init__.py:155


def resample_coeffs(in_size: int, out_size: int) -> Tuple[int, np.ndarray, np.ndarray]:
    """Pillow's ``precompute_coeffs`` + ``normalize_coeffs_8bpc`` (bicubic, box = whole axis) through the C ABI
    (host code, no GPU needed). Returns (ksize, bounds int32 [out, 2], kk int32 [out, ksize])."""
    lib = _lib.load()
    ksize = lib.aaclip_resample_ksize(int(in_size), int(out_size))
    if ksize <= 0:
        _lib.check(ksize, "aaclip_resample_ksize")
    bounds = np.zeros((out_size, 2), np.int32)
    kk = np.zeros((out_size, ksize), np.int32)
    _lib.check(lib.aaclip_resample_coeffs_host(int(in_size), int(out_size), bounds.ctypes.data_as(ctypes.c_void_p),
                                               kk.ctypes.data_as(ctypes.c_void_p), kk.size),
               "aaclip_resample_coeffs_host")
    return ksize, bounds, kk


class ImagePreprocessor:
    """``pre = ImagePreprocessor(518, device); x = pre(images_u8)`` with ``images_u8`` a uint8 tensor / array
    [B, H, W, 3] (or [H, W, 3]) of decoded RGB images of one size -> float32 [B, 3, S, S] on `device`, equal bit for
    bit to ``torch.stack([transform_x(Image.fromarray(im)) for im in images_u8])`` of the reference dataset."""

    def __init__(self, img_size: int, device=None, mean: Sequence[float] = CLIP_MEAN, std: Sequence[float] = CLIP_STD):
        self.lib = _lib.load()
        if not torch.cuda.is_available():
            raise RuntimeError("ImagePreprocessor needs a CUDA sm_100a device: there is no CPU path")
        self.device = torch.device(device if device is not None else "cuda")
        self.S = int(img_size)
        self._mean = (ctypes.c_float * 3)(*mean)
        self._std = (ctypes.c_float * 3)(*std)
        self._tables: Dict[int, Tuple[torch.Tensor, torch.Tensor]] = {}
        self._ws = None

    def _table(self, in_size: int):
        if in_size not in self._tables:
            _, bounds, kk = resample_coeffs(in_size, self.S)
            self._tables[in_size] = (torch.from_numpy(bounds).to(self.device), torch.from_numpy(kk).to(self.device))
        return self._tables[in_size]

    def __call__(self, images_u8, out: torch.Tensor = None) -> torch.Tensor:
        x = torch.as_tensor(images_u8)
        if x.dim() == 3:
            x = x[None]
        if x.dtype != torch.uint8 or x.dim() != 4 or x.shape[-1] != 3:
            raise ValueError(f"expected uint8 [B, H, W, 3] RGB images, got {x.dtype} {tuple(x.shape)}")
        x = x.to(self.device, non_blocking=True).contiguous()
        B, H, W, _ = x.shape
        xb, xk = self._table(W)
        yb, yk = self._table(H)
        need = self.lib.aaclip_preprocess_workspace_bytes(B, H, self.S)
        if self._ws is None or self._ws.numel() < need:
            self._ws = torch.empty(need, dtype=torch.uint8, device=self.device)
        if out is None:
            out = torch.empty((B, 3, self.S, self.S), dtype=torch.float32, device=self.device)
        elif out.shape != (B, 3, self.S, self.S) or out.dtype != torch.float32 or not out.is_contiguous():
            raise ValueError("out must be a contiguous float32 [B, 3, S, S] tensor")
        with torch.cuda.device(self.device):
            stream = torch.cuda.current_stream().cuda_stream
            _lib.check(self.lib.aaclip_preprocess_u8(x.data_ptr(), B, H, W, self.S, xb.data_ptr(), xk.data_ptr(),
                                                     yb.data_ptr(), yk.data_ptr(), self._mean, self._std,
                                                     self._ws.data_ptr(), self._ws.numel(), out.data_ptr(), stream),
                       "aaclip_preprocess_u8")
        return out
