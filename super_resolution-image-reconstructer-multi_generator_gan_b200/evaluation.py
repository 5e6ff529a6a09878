"""Evaluation-side pieces of the reference: checkpoint interop, ImageEnhancer, PSNR.

* ``load_reference_checkpoint``: the reference saves ``state_dict()`` of DDP-wrapped modules (keys prefixed with
  ``module.``, src/train.py:123-125) and strips the prefix when evaluating (src/evaluation.py:24-31); both spellings
  load here, into SRResNet or Discriminator.
* ``resume_learning_rates``: the ``continue_training`` protocol (src/train.py:51-59) divides both learning rates by 5.
* ``ImageEnhancer``: src/models.py:28-41 (Laplacian sharpening + clamp), one fused kernel.
* ``calculate_psnr``: src/utils.py:141-144 (skimage PSNR with data_range=1 over the whole tensor) on the device.
"""
from __future__ import annotations

import math
from collections import OrderedDict
from ctypes import c_void_p
from typing import Mapping, Tuple

import torch

from . import _lib
from ._lib import check, stream_ptr


def strip_module_prefix(state_dict: Mapping[str, torch.Tensor]) -> "OrderedDict[str, torch.Tensor]":
    out = OrderedDict()
    for k, v in state_dict.items():
        out[k[7:] if k.startswith("module.") else k] = v       # src/evaluation.py:26-29
    return out


def load_reference_checkpoint(module, checkpoint, strict: bool = True):
    """``checkpoint``: a path written by the reference's ``torch.save(model.state_dict(), ...)`` or a state dict."""
    if not isinstance(checkpoint, Mapping):
        checkpoint = torch.load(checkpoint, map_location="cpu", weights_only=True)
    return module.load_state_dict(strip_module_prefix(checkpoint), strict=strict)


def save_reference_checkpoint(module, path: str, ddp_prefix: bool = True) -> None:
    """Writes what the reference's training run writes (src/train.py:123-125): the state dict of the DDP wrapper."""
    sd = module.state_dict()
    if ddp_prefix:
        sd = OrderedDict(("module." + k, v.detach().cpu().clone()) for k, v in sd.items())
    torch.save(sd, path)


def resume_learning_rates(lr_generator: float, lr_discriminator: float) -> Tuple[float, float]:
    """continue_training=True (src/train.py:51-59): both learning rates / 5 for the "Post-Training" (GAN fine-tune) run."""
    return lr_generator / 5, lr_discriminator / 5


class ImageEnhancer:
    """Drop-in for the reference's ImageEnhancer (src/models.py:28-41): ``x + factor * Laplacian3x3(x)``, clamp [0, 1]."""

    def __init__(self, factor: float = 1):
        self.factor = factor

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        if not x.is_cuda or x.dim() != 4:
            raise RuntimeError("ImageEnhancer (libsrgan_b200): N x C x H x W CUDA tensors only")
        x = x.contiguous().float()
        out = torch.empty_like(x)
        N, C, H, W = x.shape
        check(_lib.lib().srg_image_enhance(c_void_p(x.data_ptr()), N, C, H, W, float(self.factor), c_void_p(out.data_ptr()),
                                           stream_ptr()), "srg_image_enhance")
        return out

    __call__ = forward


def calculate_psnr(img1: torch.Tensor, img2: torch.Tensor) -> float:
    """skimage.metrics.peak_signal_noise_ratio(img1, img2, data_range=1) as used at src/utils.py:141-144."""
    if img1.shape != img2.shape or not img1.is_cuda:
        raise RuntimeError("calculate_psnr: two CUDA tensors of the same shape")
    a, b = img1.contiguous().float(), img2.contiguous().float()
    L = _lib.lib()
    scratch = torch.empty(int(L.srg_recon_loss_scratch_bytes()), dtype=torch.uint8, device=a.device)
    out = torch.empty(1, dtype=torch.float64, device=a.device)
    check(L.srg_mse(c_void_p(a.data_ptr()), c_void_p(b.data_ptr()), a.numel(), c_void_p(scratch.data_ptr()), scratch.numel(),
                    c_void_p(out.data_ptr()), stream_ptr()), "srg_mse")
    mse = float(out.item())
    return float("inf") if mse == 0 else 10.0 * math.log10(1.0 / mse)
