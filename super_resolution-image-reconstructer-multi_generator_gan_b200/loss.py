"""Losses of the reference's train step over libsrgan_b200.so.

``ReconstructionLoss`` mirrors src/utils.py:173-241 (called as ``g_criterion(hr_imgs, sr_images)`` at
src/train.py:189): returns ``(edge_loss, tv_loss)``, differentiable w.r.t. the second argument only (the reference's
edge weights depend on the HR image, which carries no gradient).  ``tanh_mean`` is the relativistic term of
src/train.py:190 / :218.  Both are fused, vectorised reduction kernels; no PyTorch fallback.
"""
from __future__ import annotations

from ctypes import c_void_p

import torch
import torch.nn as nn

from . import _lib
from ._lib import check, stream_ptr


def _scratch(device) -> torch.Tensor:
    n = int(_lib.lib().srg_recon_loss_scratch_bytes())
    return torch.empty(n, dtype=torch.uint8, device=device)


def _as_f32_cuda(t: torch.Tensor, what: str) -> torch.Tensor:
    if not t.is_cuda:
        raise RuntimeError(f"{what}: CUDA tensors only (libsrgan_b200 has no CPU path)")
    t = t.contiguous()
    return t if t.dtype == torch.float32 else t.float()


class _ReconLossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, hr, sr):
        L = _lib.lib()
        N, C, H, W = sr.shape
        scratch = _scratch(sr.device)
        e_buf = torch.empty_like(sr)
        g_buf = torch.empty_like(sr)
        losses = torch.empty(2, dtype=torch.float32, device=sr.device)
        check(L.srg_recon_loss_forward(c_void_p(hr.data_ptr()), c_void_p(sr.data_ptr()), N, C, H, W,
                                       c_void_p(scratch.data_ptr()), scratch.numel(), c_void_p(e_buf.data_ptr()),
                                       c_void_p(g_buf.data_ptr()), c_void_p(losses.data_ptr()), stream_ptr()),
              "srg_recon_loss_forward")
        ctx.save_for_backward(hr, sr, scratch, e_buf, g_buf)
        ctx.set_materialize_grads(False)
        edge, tv = losses[0], losses[1]
        return edge, tv

    @staticmethod
    def backward(ctx, g_edge, g_tv):
        hr, sr, scratch, e_buf, g_buf = ctx.saved_tensors
        if g_edge is None and g_tv is None:
            return None, None
        L = _lib.lib()
        N, C, H, W = sr.shape
        zero = None
        if g_edge is None or g_tv is None:
            zero = torch.zeros((), dtype=torch.float32, device=sr.device)
        w_e = (g_edge if g_edge is not None else zero).contiguous().float()
        w_t = (g_tv if g_tv is not None else zero).contiguous().float()
        grad = torch.empty_like(sr)
        check(L.srg_recon_loss_backward(c_void_p(hr.data_ptr()), c_void_p(sr.data_ptr()), N, C, H, W,
                                        c_void_p(scratch.data_ptr()), c_void_p(e_buf.data_ptr()),
                                        c_void_p(g_buf.data_ptr()), c_void_p(w_e.data_ptr()), c_void_p(w_t.data_ptr()),
                                        c_void_p(grad.data_ptr()), 1.0, stream_ptr()), "srg_recon_loss_backward")
        return None, grad


class ReconstructionLoss(nn.Module):
    """Edge-weighted L1 + Laplacian TV penalty (reference: src/utils.py:173-241, closed form in SURVEY Appendix C).

    ``forward(original_images=HR, target_images=SR) -> (edge_loss, tv_loss)``; statistics (mean / unbiased std of the
    edge map, sum of edge weights) are over the whole local batch, exactly like the reference (per rank under DDP).
    """

    def __init__(self):
        super().__init__()

    def forward(self, original_images: torch.Tensor, target_images: torch.Tensor):
        if original_images.shape != target_images.shape or target_images.dim() != 4:
            raise RuntimeError(f"ReconstructionLoss: shape mismatch {tuple(original_images.shape)} vs {tuple(target_images.shape)}")
        hr = _as_f32_cuda(original_images.detach(), "ReconstructionLoss")
        sr = _as_f32_cuda(target_images, "ReconstructionLoss")
        return _ReconLossFn.apply(hr, sr)


class _TanhMeanFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, a, b, sign):
        L = _lib.lib()
        scratch = _scratch(a.device)
        out = torch.empty(1, dtype=torch.float32, device=a.device)
        need_a, need_b = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        da = torch.empty_like(a) if need_a else None
        db = torch.empty_like(b) if need_b else None
        check(L.srg_tanh_mean(c_void_p(a.data_ptr()), c_void_p(b.data_ptr()), a.numel(), float(sign),
                              c_void_p(scratch.data_ptr()), scratch.numel(), c_void_p(out.data_ptr()),
                              c_void_p(da.data_ptr()) if da is not None else None,
                              c_void_p(db.data_ptr()) if db is not None else None, 1.0, stream_ptr()), "srg_tanh_mean")
        ctx.da, ctx.db = da, db
        return out[0]

    @staticmethod
    def backward(ctx, g):
        da, db = ctx.da, ctx.db
        ctx.da = ctx.db = None
        # g is the (scalar) upstream gradient: scale the stored unit gradients in place (tiny tensors: N x 512 x h x w)
        if da is not None:
            da.mul_(g)
        if db is not None:
            db.mul_(g)
        return da, db, None


def tanh_mean(a: torch.Tensor, b: torch.Tensor, sign: float = 1.0) -> torch.Tensor:
    """mean(tanh(sign * (a - b))): ``tanh_mean(fake, real)`` is the discriminator loss of src/train.py:218,
    ``tanh_mean(real, fake)`` the generator's adversarial term of src/train.py:190."""
    if a.shape != b.shape:
        raise RuntimeError("tanh_mean: shape mismatch")
    return _TanhMeanFn.apply(_as_f32_cuda(a, "tanh_mean"), _as_f32_cuda(b, "tanh_mean"), sign)


class _PointLossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, a, b, kind):
        L = _lib.lib()
        scratch = _scratch(a.device)
        out = torch.empty(1, dtype=torch.float32, device=a.device)
        ga = torch.empty_like(a) if ctx.needs_input_grad[0] else None
        check(L.srg_point_loss(int(kind), c_void_p(a.data_ptr()), c_void_p(b.data_ptr()), a.numel(),
                               c_void_p(scratch.data_ptr()), scratch.numel(), c_void_p(out.data_ptr()),
                               c_void_p(ga.data_ptr()) if ga is not None else None, 1.0, stream_ptr()), "srg_point_loss")
        ctx.ga = ga
        return out[0]

    @staticmethod
    def backward(ctx, g):
        ga = ctx.ga
        ctx.ga = None
        if ga is not None:
            ga.mul_(g)
        return ga, None, None


def _point(a, b, kind, what):
    if a.shape != b.shape:
        raise RuntimeError(f"{what}: shape mismatch")
    return _PointLossFn.apply(_as_f32_cuda(a, what), _as_f32_cuda(b.detach(), what), kind)


def l1_loss(input: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
    """mean |input - target| (torch.nn.functional.l1_loss); gradient w.r.t. ``input``."""
    return _point(input, target, 0, "l1_loss")


def mse_loss(input: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
    """mean (input - target)^2 (torch.nn.functional.mse_loss); gradient w.r.t. ``input``."""
    return _point(input, target, 1, "mse_loss")


def bce_loss(probabilities: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
    """binary cross entropy on probabilities, e.g. the discriminator's sigmoid map against 1 / 0 labels
    (torch.nn.functional.binary_cross_entropy, logs clamped at -100); gradient w.r.t. ``probabilities``."""
    return _point(probabilities, target, 2, "bce_loss")
