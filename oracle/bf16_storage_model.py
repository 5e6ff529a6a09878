"""bf16-STORAGE model of the generator arithmetic.  TEST INFRASTRUCTURE ONLY (same rules as srgan_oracle.py).

The CUDA path keeps activations and inter-layer gradients in bf16 and multiplies bf16 operands with fp32
accumulation.  Against the fp32 oracle that costs ~2e-3 relative error per layer, and -- because ReLU masks are
taken on slightly different values -- a few percent on deep-layer gradients, which says nothing about kernel
correctness.  This model is the oracle's SRResNet (src/models.py:44-87, :10-25) with a round-to-bf16 inserted at
exactly the points where the CUDA engine stores a tensor, so kernel bugs (indexing, taps, masks, reductions) show up
as errors far above the remaining fp32-accumulation-order noise (~1e-3).  It is a diagnostic companion to the fp32
oracle, never a replacement: parity claims are made against srgan_oracle.py / the golden fixtures.
"""
from __future__ import annotations

from typing import Dict, Optional

import torch
import torch.nn.functional as F

from . import srgan_oracle as O

Tensor = torch.Tensor


def _r(x: Tensor) -> Tensor:
    return x.to(torch.bfloat16).to(torch.float32)


class _RoundFwd(torch.autograd.Function):
    """value stored in bf16 (forward); gradient passes unchanged"""
    @staticmethod
    def forward(ctx, x):
        return _r(x)

    @staticmethod
    def backward(ctx, g):
        return g


class _RoundGrad(torch.autograd.Function):
    """identity forward; the gradient arriving here is stored in bf16"""
    @staticmethod
    def forward(ctx, x):
        return x.view_as(x)

    @staticmethod
    def backward(ctx, g):
        return _r(g)


def rf(x):
    return _RoundFwd.apply(x)


def rg(x):
    return _RoundGrad.apply(x)


def _conv(x, w, b, pad):
    # bf16 operands, fp32 accumulation; the straight-through weight rounding leaves d/dw defined on the fp32 master
    wq = w + (_r(w.detach()) - w.detach())
    return F.conv2d(x, wq, b, padding=pad)


def _bn_train(y: Tensor, sd, prefix: str) -> Tensor:
    mean = y.mean(dim=(0, 2, 3))
    var = y.var(dim=(0, 2, 3), unbiased=False)
    inv = torch.rsqrt(var + O.BN_EPS)
    return (y - mean[None, :, None, None]) * (inv * sd[prefix + ".weight"])[None, :, None, None] + \
        sd[prefix + ".bias"][None, :, None, None]


def srresnet_forward_train(sd: Dict[str, Tensor], x: Tensor, taps: Optional[Dict[str, Tensor]] = None) -> Tensor:
    def tap(n, t):
        if taps is not None:
            taps[n] = t
        return t

    x = _r(x)                                                   # unfold9 stores the LR image in bf16
    pre1 = rg(_conv(x, sd["conv1.weight"], sd["conv1.bias"], 4))   # gradient w.r.t. the pre-activation is stored
    out1 = rf(F.leaky_relu(pre1, 0.2))
    tap("out1", out1)
    out = out1
    n_blocks = len({k.split(".")[1] for k in sd if k.startswith("residual_blocks.") and k.endswith("conv1.weight")})
    for i in range(n_blocks):
        p = f"residual_blocks.{i}"
        xin = rg(out)                                           # d(x_in) = dgrad + skip gradient, stored
        y1 = rg(rf(_conv(xin, sd[p + ".conv1.weight"], sd[p + ".conv1.bias"], 1)))
        tap(p + ".conv1", y1)
        z1 = rf(F.relu(rg(_bn_train(y1, sd, p + ".bn1"))))
        y2 = rg(rf(_conv(z1, sd[p + ".conv2.weight"], sd[p + ".conv2.bias"], 1)))
        tap(p + ".conv2", y2)
        out = rf(_bn_train(y2, sd, p + ".bn2") + xin)
        tap(p, out)
    out = rg(out)
    trunk = rg(rf(_conv(out, sd["conv2.weight"], sd["conv2.bias"], 1) + out1))
    tap("trunk", trunk)
    out = trunk
    for j in O.upsample_stage_indices(sd):
        pre = rg(F.pixel_shuffle(_conv(out, sd[f"upsample.{j}.weight"], sd[f"upsample.{j}.bias"], 1), 2))
        out = rf(F.relu(pre))
        tap(f"upsample.{j}", out)
    sr = _conv(out, sd["conv3.weight"], sd["conv3.bias"], 4)
    return sr
