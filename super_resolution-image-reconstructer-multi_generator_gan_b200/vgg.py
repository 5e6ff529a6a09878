"""VGG19 perceptual loss of the reference over libsrgan_b200.so.

``VGGFeatureExtractor`` mirrors src/models.py:123-151: torchvision's ``vgg19.features`` as ``self.vgg19`` (same
``state_dict`` keys ``vgg19.<i>.weight / .bias``, frozen parameters), ``layer_name_mapping`` {'3', '8', '17', '26', '35'}
and ``selected_layers`` (default ('conv3_3', 'conv4_3') = the ReLU outputs at indices 17 and 26), early exit once every
selected layer is collected.  ``perceptal_loss`` [sic] mirrors src/utils.py:154-166: sum over the selected layers of
``L1Loss(features(sr), features(hr))``, differentiable w.r.t. ``sr_imgs`` only.

The reference downloads ImageNet weights (``VGG19_Weights.DEFAULT``); there is no network here, so the module starts
from torchvision's default initialisation (kaiming-normal, fan_out) and accepts the pretrained tensors through
``load_state_dict`` / ``load_torchvision_features`` when a user has them.

Execution: every conv + ReLU is one ``srg_conv2d_fprop`` launch (tcgen05 implicit GEMM, bias + ReLU in the epilogue) on
NHWC bf16 activations; the 3-channel first layer is an im2col pass + a 1x1 convolution; MaxPool2d(2, 2), the L1 reduction
and the whole backward chain (``srg_conv2d_dgrad`` with the ReLU mask in the epilogue, pool backward, col2im) are the
library's kernels as well.  No PyTorch / cuDNN fallback.
"""
from __future__ import annotations

import math
from ctypes import c_void_p
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn as nn

from . import _lib
from ._lib import check, stream_ptr

# torchvision vgg19 configuration 'E' (features only): channel counts, 'M' = MaxPool2d(2, 2)
_VGG19_CFG = [64, 64, "M", 128, 128, "M", 256, 256, 256, 256, "M", 512, 512, 512, 512, "M", 512, 512, 512, 512, "M"]


class _VggConv(nn.Module):
    """Parameter holder with nn.Conv2d's attribute names (weight OIHW, bias); executed by the extractor."""

    def __init__(self, cin: int, cout: int):
        super().__init__()
        self.in_channels, self.out_channels = cin, cout
        self.weight = nn.Parameter(torch.empty(cout, cin, 3, 3))
        self.bias = nn.Parameter(torch.zeros(cout))
        nn.init.kaiming_normal_(self.weight, mode="fan_out", nonlinearity="relu")      # torchvision/models/vgg.py

    def forward(self, x):
        raise RuntimeError("VGG layers run inside VGGFeatureExtractor (libsrgan_b200); they are not callable on their own")


class _VggMarker(nn.Module):
    def __init__(self, kind: str):
        super().__init__()
        self.kind = kind

    def forward(self, x):
        raise RuntimeError("VGG layers run inside VGGFeatureExtractor (libsrgan_b200); they are not callable on their own")

    def extra_repr(self):
        return self.kind


def _p(t: torch.Tensor) -> c_void_p:
    return c_void_p(t.data_ptr())


class _Plan:
    """Layer walk of one forward: list of ('conv', idx, cin, cout) / ('pool', idx) up to the last selected layer."""

    def __init__(self, modules, mapping: Dict[str, str], selected: Sequence[str]):
        self.steps: List[Tuple] = []
        self.taps: Dict[int, str] = {}          # feature index (module index of the ReLU) -> layer name
        want = [n for n in mapping.values() if n in selected]
        found = 0
        i = 0
        mods = list(modules)
        while i < len(mods):
            m = mods[i]
            if isinstance(m, _VggConv):
                self.steps.append(("conv", i, m.in_channels, m.out_channels))
                i += 1                           # the ReLU that follows is fused into the conv epilogue
                if i >= len(mods) or getattr(mods[i], "kind", "") != "ReLU":
                    raise RuntimeError("VGGFeatureExtractor expects conv -> ReLU pairs")
            elif getattr(m, "kind", "") == "MaxPool2d":
                self.steps.append(("pool", i))
            name = mapping.get(str(i))
            if name is not None and name in selected:
                self.taps[len(self.steps) - 1] = name
                found += 1
                if found == len(want):
                    break
            i += 1
        if found != len(want):
            raise RuntimeError("selected layers not reachable")


class VGGFeatureExtractor(nn.Module):
    """Drop-in for the reference's VGGFeatureExtractor (src/models.py:123-151)."""

    def __init__(self, layers=("conv3_3", "conv4_3")):
        super().__init__()
        mods: List[nn.Module] = []
        cin = 3
        for v in _VGG19_CFG:
            if v == "M":
                mods.append(_VggMarker("MaxPool2d"))
            else:
                mods += [_VggConv(cin, int(v)), _VggMarker("ReLU")]
                cin = int(v)
        self.layer_name_mapping = {"3": "conv1_2", "8": "conv2_2", "17": "conv3_3", "26": "conv4_3", "35": "conv5_3"}
        self.selected_layers = layers
        self.vgg19 = nn.Sequential(*mods)
        for p in self.vgg19.parameters():
            p.requires_grad = False
        self._packed = None

    # ---- weights ---------------------------------------------------------------------------------------------------
    def load_torchvision_features(self, features: nn.Module) -> None:
        """Copy the tensors of a torchvision ``vgg19().features`` (e.g. one holding the pretrained weights)."""
        self.vgg19.load_state_dict(features.state_dict())
        self._packed = None

    def load_state_dict(self, *a, **k):
        self._packed = None
        return super().load_state_dict(*a, **k)

    def _apply(self, fn, *a, **k):
        self._packed = None
        return super()._apply(fn, *a, **k)

    def _pack(self, plan: _Plan, device):
        """bf16 operand copies of every executed conv: forward filter, transposed / flipped filter for the input gradient."""
        if self._packed is not None and self._packed[0] == device:
            return self._packed[1]
        L = _lib.lib()
        packed = {}
        for st in plan.steps:
            if st[0] != "conv":
                continue
            _, idx, cin, cout = st
            m = self.vgg19[idx]
            w = m.weight.detach().to(device=device, dtype=torch.float32).contiguous()
            if cin == 3:
                # first layer as a 1x1 convolution over the im2col channels (kh*3 + kw)*3 + c
                w1 = torch.zeros(cout, 64, 1, 1, device=device)
                w1[:, :27, 0, 0] = w.permute(0, 2, 3, 1).reshape(cout, 27)
                w, k, kc = w1, 1, 64
            else:
                k, kc = 3, cin
            nbytes = int(L.srg_conv2d_packed_weight_bytes(cout, kc, k))
            wf = torch.empty(nbytes, dtype=torch.uint8, device=device)
            wd = torch.empty(nbytes, dtype=torch.uint8, device=device)
            check(L.srg_conv2d_pack_weights(_p(w), cout, kc, k, 0, _p(wf), stream_ptr()), "srg_conv2d_pack_weights")
            check(L.srg_conv2d_pack_weights(_p(w), cout, kc, k, 1, _p(wd), stream_ptr()), "srg_conv2d_pack_weights")
            packed[idx] = (wf, wd, m.bias.detach().to(device=device, dtype=torch.float32).contiguous(), k, kc)
        self._packed = (device, packed)
        return packed

    def _plan(self) -> _Plan:
        return _Plan(self.vgg19, self.layer_name_mapping, self.selected_layers)

    # ---- execution -------------------------------------------------------------------------------------------------
    def _run(self, x: torch.Tensor, keep_all: bool):
        """Forward through the plan.  Returns (activations per step [NHWC bf16], geometry per step, unfolded input)."""
        if not x.is_cuda:
            raise RuntimeError("VGGFeatureExtractor (libsrgan_b200) runs on a CUDA device only; there is no CPU path")
        if x.dim() != 4 or x.shape[1] != 3:
            raise RuntimeError(f"expected an N x 3 x H x W input, got {tuple(x.shape)}")
        x = x.detach().contiguous().float()
        N, _, H, W = x.shape
        plan = self._plan()
        packed = self._pack(plan, x.device)
        L = _lib.lib()
        acts: List[Optional[torch.Tensor]] = []
        geo: List[Tuple[int, int, int]] = []
        cur = torch.empty(N, H, W, 64, dtype=torch.bfloat16, device=x.device)
        check(L.srg_unfold3x3_rgb(_p(x), N, H, W, _p(cur), stream_ptr()), "srg_unfold3x3_rgb")
        unf = cur
        h, w = H, W
        last_tap = max(plan.taps)
        for si, st in enumerate(plan.steps):
            if st[0] == "conv":
                _, idx, cin, cout = st
                wf, _, bias, k, kc = packed[idx]
                out = torch.empty(N, h, w, cout, dtype=torch.bfloat16, device=x.device)
                check(L.srg_conv2d_fprop(_p(cur), N, h, w, kc, _p(wf), cout, k, _p(bias), 1, 0.0, None, _p(out), stream_ptr()),
                      "srg_conv2d_fprop")
                c = cout
            else:
                c = cur.shape[3]
                if h < 2 or w < 2:
                    raise RuntimeError("VGGFeatureExtractor: input too small for the selected layers")
                out = torch.empty(N, h // 2, w // 2, c, dtype=torch.bfloat16, device=x.device)
                check(L.srg_maxpool2x2_forward(_p(cur), N, h, w, c, _p(out), stream_ptr()), "srg_maxpool2x2_forward")
                h, w = h // 2, w // 2
            acts.append(out if (keep_all or si in plan.taps) else None)
            geo.append((h, w, c))
            cur = out
            if si == last_tap:
                break
        return plan, acts, geo, unf

    def forward(self, x: torch.Tensor) -> Dict[str, torch.Tensor]:
        """{layer name: N x C x h x w fp32 feature map} like the reference (no autograd graph: the parameters are frozen
        and the differentiable path through the input is ``perceptal_loss``)."""
        plan, acts, _, _ = self._run(x, keep_all=False)
        return {name: acts[si].float().permute(0, 3, 1, 2).contiguous() for si, name in plan.taps.items()}


class _PerceptualFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, sr, hr, fe: VGGFeatureExtractor):
        L = _lib.lib()
        dev = sr.device
        plan, real, _, _ = fe._run(hr, keep_all=False)
        plan, fake, geo, unf = fe._run(sr, keep_all=True)
        N = sr.shape[0]
        loss = torch.zeros(1, dtype=torch.float32, device=dev)
        scratch = torch.empty(int(L.srg_l1_bf16_scratch_bytes()), dtype=torch.uint8, device=dev)
        l1_grads = {}
        for si in sorted(plan.taps):
            a, b = fake[si], real[si]
            g = torch.empty_like(a)
            # gradient of this layer's L1 term w.r.t. the (ReLU) feature, already masked by the ReLU: d loss / d pre-activation
            check(L.srg_l1_bf16(_p(a), _p(b), a.numel(), 1.0, 1, 1.0, 1, _p(g), _p(scratch), scratch.numel(), _p(loss),
                                stream_ptr()), "srg_l1_bf16")
            l1_grads[si] = g
        ctx.fe, ctx.plan, ctx.fake, ctx.geo, ctx.unf, ctx.l1_grads = fe, plan, fake, geo, unf, l1_grads
        ctx.shape = tuple(sr.shape)
        return loss[0]

    @staticmethod
    def backward(ctx, g_out):
        fe, plan, fake, geo, unf, l1_grads = ctx.fe, ctx.plan, ctx.fake, ctx.geo, ctx.unf, ctx.l1_grads
        L = _lib.lib()
        N, _, H, W = ctx.shape
        dev = unf.device
        packed = fe._pack(plan, dev)
        last = max(plan.taps)
        # dz = d loss / d (pre-ReLU output of step si); walk the plan backwards
        dz = l1_grads[last]
        for si in range(last, -1, -1):
            st = plan.steps[si]
            h_in, w_in, c_in = (geo[si - 1] if si > 0 else (H, W, 64))
            x_in = fake[si - 1] if si > 0 else unf
            if st[0] == "conv":
                _, idx, cin, cout = st
                _, wd, _, k, kc = packed[idx]
                dx = torch.empty(N, h_in, w_in, kc, dtype=torch.bfloat16, device=dev)
                # input gradient, masked by the ReLU that produced x_in (a pooled map is a max of ReLU outputs: same mask);
                # the unfolded image (si == 0) has no activation in front of it
                mask = _p(x_in) if si > 0 else None
                h_o, w_o, _ = geo[si]
                check(L.srg_conv2d_dgrad(_p(dz), N, h_o, w_o, cout, _p(wd), kc, k, mask, None, _p(dx), stream_ptr()),
                      "srg_conv2d_dgrad")
                dz = dx
                if si - 1 in l1_grads:           # a feature tapped right below this conv (not the case for VGG19's taps)
                    dz = dz + l1_grads[si - 1]
            else:
                add = l1_grads.get(si - 1)
                dx = torch.empty(N, h_in, w_in, c_in, dtype=torch.bfloat16, device=dev)
                check(L.srg_maxpool2x2_backward(_p(x_in), _p(dz), _p(add) if add is not None else None, N, h_in, w_in, c_in,
                                                _p(dx), stream_ptr()), "srg_maxpool2x2_backward")
                dz = dx
        d_sr = torch.empty(N, 3, H, W, dtype=torch.float32, device=dev)
        check(L.srg_fold3x3_rgb(_p(dz), N, H, W, 1.0, _p(d_sr), stream_ptr()), "srg_fold3x3_rgb")
        if g_out is not None:
            d_sr = d_sr * g_out
        return d_sr, None, None


def perceptal_loss(sr_imgs: torch.Tensor, hr_imgs: torch.Tensor, feature_extractor: VGGFeatureExtractor) -> torch.Tensor:
    """src/utils.py:154-166 (name as upstream): sum over the selected VGG layers of L1Loss(features(sr), features(hr))."""
    if sr_imgs.shape != hr_imgs.shape:
        raise RuntimeError(f"perceptal_loss: shape mismatch {tuple(sr_imgs.shape)} vs {tuple(hr_imgs.shape)}")
    return _PerceptualFn.apply(sr_imgs, hr_imgs, feature_extractor)


perceptual_loss = perceptal_loss
