"""On-GPU counterparts of the reference's per-image PIL transforms (src/transformers.py:73-94, applied in the dataset's
``__getitem__``, src/utils.py:34-47) for BATCHES of 8-bit RGB images already on the device.

Reference: ``downward_img_quality`` = Resize((clip_height // 4, clip_width // 4)) [PIL, antialiased bilinear] -> ToTensor
-> ``x + randn_like(x) * uniform(0, 0.03)``; ``normalize_img_size`` = Resize((clip_height, clip_width), BICUBIC) ->
ToTensor; ``to_tensor``; ``add_noise``.  With ``num_workers=0`` PIL loading (src/train.py:94-95) these transforms starve
8 GPUs; here one launch pair per batch does the resize bit-exactly like Pillow (``srg_resize_u8``) with ToTensor and the
degradation fused into the second pass.  Random numbers come from torch's device generator (plumbing): the noise field is
``torch.randn`` and the per-image sigma ``torch.rand * 0.03``, so values are not those of the reference's CPU / Python
RNG streams -- their distribution is.

Inputs: uint8 CUDA tensors ``[N, H, W, 3]`` (decoded RGB, channel-last like PIL).  No CPU path.
"""
from __future__ import annotations

from ctypes import c_void_p
from typing import Dict, Optional, Tuple

import numpy as np
import torch

from . import _lib
from ._lib import check, stream_ptr

BILINEAR, BICUBIC = 0, 1
clip_width, clip_height = 1024, 512          # src/variables.py:5-6

_plans: Dict[Tuple[int, int, int, str], Tuple[torch.Tensor, torch.Tensor, int]] = {}


def _plan(in_size: int, out_size: int, filt: int, device) -> Tuple[torch.Tensor, torch.Tensor, int]:
    key = (in_size, out_size, filt, str(device))
    if key not in _plans:
        L = _lib.lib()
        k = int(L.srg_resize_plan_ksize(in_size, out_size, filt))
        if k < 1:
            raise RuntimeError(f"resize: bad sizes {in_size} -> {out_size} or filter {filt}")
        bounds = np.zeros((out_size, 2), dtype=np.int32)
        coeffs = np.zeros((out_size, k), dtype=np.int32)
        check(L.srg_resize_plan(in_size, out_size, filt, c_void_p(bounds.ctypes.data), c_void_p(coeffs.ctypes.data)),
              "srg_resize_plan")
        _plans[key] = (torch.from_numpy(bounds).to(device), torch.from_numpy(coeffs).to(device), k)
    return _plans[key]


def _check_u8(img: torch.Tensor) -> torch.Tensor:
    if not img.is_cuda:
        raise RuntimeError("transformers (libsrgan_b200): CUDA tensors only; there is no CPU path")
    if img.dtype != torch.uint8 or img.dim() != 4 or img.shape[3] != 3:
        raise RuntimeError(f"expected a uint8 [N, H, W, 3] batch, got {img.dtype} {tuple(img.shape)}")
    return img.contiguous()


def _resize(img: torch.Tensor, out_h: int, out_w: int, filt: int, want_u8: bool, want_f32: bool,
            noise: Optional[torch.Tensor] = None, sigma: Optional[torch.Tensor] = None):
    img = _check_u8(img)
    N, H, W, _ = img.shape
    dev = img.device
    bw, cw, kw = _plan(W, out_w, filt, dev) if out_w != W else (None, None, 0)
    bh, ch, kh = _plan(H, out_h, filt, dev)
    tmp = torch.empty(N, H, out_w, 3, dtype=torch.uint8, device=dev) if out_w != W else None
    out_u8 = torch.empty(N, out_h, out_w, 3, dtype=torch.uint8, device=dev) if want_u8 else None
    out_f = torch.empty(N, 3, out_h, out_w, dtype=torch.float32, device=dev) if want_f32 else None
    if noise is not None:
        noise = noise.contiguous().float()
        sigma = sigma.contiguous().float()
        if tuple(noise.shape) != (N, 3, out_h, out_w) or sigma.numel() != N:
            raise RuntimeError("resize: noise must be [N, 3, out_h, out_w] and sigma [N]")

    def p(t):
        return c_void_p(t.data_ptr()) if t is not None else None

    check(_lib.lib().srg_resize_u8(p(img), N, H, W, out_h, out_w, p(bw), p(cw), kw, p(bh), p(ch), kh, p(tmp), p(out_u8), p(out_f),
                                   p(noise), p(sigma), stream_ptr()), "srg_resize_u8")
    return out_u8, out_f


def resize_u8(img: torch.Tensor, size: Tuple[int, int], interpolation: int = BILINEAR) -> torch.Tensor:
    """``PIL.Image.resize`` (what ``transforms.Resize(size, interpolation)`` does to a PIL image): uint8 [N, h, w, 3]."""
    return _resize(img, int(size[0]), int(size[1]), interpolation, True, False)[0]


def to_tensor(img: torch.Tensor) -> torch.Tensor:
    """``transforms.ToTensor`` (src/transformers.py:88-90): uint8 [N, H, W, 3] -> fp32 [N, 3, H, W] in [0, 1]."""
    img = _check_u8(img)
    return _resize(img, img.shape[1], img.shape[2], BILINEAR, False, True)[1]


def normalize_img_size(img: torch.Tensor, height: int = clip_height, width: int = clip_width) -> torch.Tensor:
    """src/transformers.py:79-82: bicubic resize to (clip_height, clip_width), then ToTensor."""
    return _resize(img, height, width, BICUBIC, False, True)[1]


def downward_img_quality(img: torch.Tensor, height: int = clip_height // 4, width: int = clip_width // 4,
                         max_sigma: float = 0.03, generator: Optional[torch.Generator] = None,
                         noise: Optional[torch.Tensor] = None, sigma: Optional[torch.Tensor] = None) -> torch.Tensor:
    """src/transformers.py:73-77: antialiased bilinear resize to a quarter of the clip size, ToTensor, plus Gaussian noise
    whose standard deviation is drawn per image from U(0, max_sigma).  ``noise`` / ``sigma`` override the random draws."""
    img = _check_u8(img)
    N = img.shape[0]
    if noise is None:
        noise = torch.randn(N, 3, height, width, device=img.device, generator=generator)
    if sigma is None:
        sigma = torch.rand(N, device=img.device, generator=generator) * max_sigma
    return _resize(img, height, width, BILINEAR, False, True, noise, sigma)[1]


def add_noise(img: torch.Tensor, max_sigma: float = 0.03, generator: Optional[torch.Generator] = None,
              noise: Optional[torch.Tensor] = None, sigma: Optional[torch.Tensor] = None) -> torch.Tensor:
    """src/transformers.py:92-95: ToTensor, then the same additive noise (no resize)."""
    img = _check_u8(img)
    return downward_img_quality(img, img.shape[1], img.shape[2], max_sigma, generator, noise, sigma)


def synthesize_pair(img: torch.Tensor, generator: Optional[torch.Generator] = None) -> Tuple[torch.Tensor, torch.Tensor]:
    """(hr, lr) of one training batch from decoded uint8 images: what ImageDatasetWithTransforms.__getitem__ returns per
    image (src/utils.py:41-47: ``(norm_transform(image), quality_transform(image))``), for a whole batch on the device."""
    return normalize_img_size(img), downward_img_quality(img, generator=generator)
