"""CPU oracle for the SR-GAN hot path.  TEST INFRASTRUCTURE ONLY.

This file restates, as plain functions over a ``state_dict``-style mapping of fp32 CPU tensors, the
arithmetic of the reference's hot path (``/root/reference/src``).  Only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may
import it, and only as the checker or the CPU baseline -- never from the product package.

Where the arithmetic lives: the reference delegates every FLOP to PyTorch (``requirements.txt:5`` pins
torch 2.6.0; this image has 2.11.0).  The restatement therefore also evaluates on torch's CPU fp32
kernels, but it is written from the algorithm (state_dict keys + layer maths), not from the module code.

Parity pin: the reference has no tests or golden vectors of its own, so the pins
are outputs of the unmodified reference run in the build container: ``tests/golden/*.npz|json`` made by
``tests/golden/make_golden.py`` (committed) and re-checked by ``tests/test_oracle_golden.py``.

Each function cites the reference lines it follows.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor
BN_EPS = 1e-5
BN_MOMENTUM = 0.1


# --------------------------------------------------------------------------------------------------
# Generator: src/models.py:44-87 (SRResNet) and :10-25 (ResidualBlock)
# --------------------------------------------------------------------------------------------------
def _batch_norm(x: Tensor, sd: Dict[str, Tensor], prefix: str, training: bool, update_running: bool) -> Tensor:
    """nn.BatchNorm2d(64) semantics (src/models.py:16,19): batch statistics + biased variance when
    training, running statistics in eval; running update uses momentum 0.1 and the unbiased variance."""
    w, b = sd[prefix + ".weight"], sd[prefix + ".bias"]
    if training:
        mean = x.mean(dim=(0, 2, 3))
        var = x.var(dim=(0, 2, 3), unbiased=False)
        if update_running:
            n = x.numel() // x.shape[1]
            with torch.no_grad():
                rm, rv = sd[prefix + ".running_mean"], sd[prefix + ".running_var"]
                rm.mul_(1 - BN_MOMENTUM).add_(mean.detach(), alpha=BN_MOMENTUM)
                rv.mul_(1 - BN_MOMENTUM).add_(var.detach() * (n / max(n - 1, 1)), alpha=BN_MOMENTUM)
                key = prefix + ".num_batches_tracked"
                if key in sd:
                    sd[key] += 1
    else:
        mean, var = sd[prefix + ".running_mean"], sd[prefix + ".running_var"]
    inv = torch.rsqrt(var + BN_EPS)
    return (x - mean[None, :, None, None]) * (inv * w)[None, :, None, None] + b[None, :, None, None]


def upsample_stage_indices(sd: Dict[str, Tensor]) -> List[int]:
    """Indices of the conv layers inside ``upsample`` (0, 3, 6, ...: conv, PixelShuffle, ReLU triples,
    src/models.py:69-75; ``int(upscale_factor // 2)`` stages)."""
    idx = sorted({int(k.split(".")[1]) for k in sd if k.startswith("upsample.") and k.endswith(".weight")})
    return idx


def srresnet_forward(sd: Dict[str, Tensor], x: Tensor, training: bool, update_running: bool = True,
                     taps: Optional[Dict[str, Tensor]] = None) -> Tensor:
    """SRResNet.forward (src/models.py:80-87).  ``taps`` (optional dict) receives named intermediates
    for per-layer parity checks."""
    def tap(name, t):
        if taps is not None:
            taps[name] = t
        return t

    out1 = F.leaky_relu(F.conv2d(x, sd["conv1.weight"], sd["conv1.bias"], padding=4), 0.2)  # :81 (in-place act)
    tap("out1", out1)
    out = out1
    n_blocks = len({k.split(".")[1] for k in sd if k.startswith("residual_blocks.") and k.endswith("conv1.weight")})
    for i in range(n_blocks):  # :82, block body :21-25
        p = f"residual_blocks.{i}"
        y = F.conv2d(out, sd[p + ".conv1.weight"], sd[p + ".conv1.bias"], padding=1)
        tap(p + ".conv1", y)
        y = F.relu(_batch_norm(y, sd, p + ".bn1", training, update_running))
        y = F.conv2d(y, sd[p + ".conv2.weight"], sd[p + ".conv2.bias"], padding=1)
        tap(p + ".conv2", y)
        y = _batch_norm(y, sd, p + ".bn2", training, update_running)
        out = tap(p, y + out)
    out = F.conv2d(out, sd["conv2.weight"], sd["conv2.bias"], padding=1) + out1  # :83-84
    tap("trunk", out)
    for j in upsample_stage_indices(sd):  # :85
        out = F.conv2d(out, sd[f"upsample.{j}.weight"], sd[f"upsample.{j}.bias"], padding=1)
        out = F.relu(F.pixel_shuffle(out, 2))
        tap(f"upsample.{j}", out)
    return F.conv2d(out, sd["conv3.weight"], sd["conv3.bias"], padding=4)  # :86, no output activation


# --------------------------------------------------------------------------------------------------
# Discriminator: src/models.py:90-120
# --------------------------------------------------------------------------------------------------
def discriminator_output_hw(h: int, w: int) -> Tuple[int, int]:
    """Spatial size after the four (conv s2, MaxPool(3,2)) stages; raises like the reference would
     when the map vanishes."""
    def axis(a: int) -> int:
        a = (a + 2 * 2 - 8) // 2 + 1          # conv 8x8 s2 p2
        for stage in range(4):
            if stage > 0:
                a = (a + 2 * 1 - 4) // 2 + 1  # conv 4x4 s2 p1
            if a < 3:
                raise RuntimeError("Discriminator input too small: MaxPool2d(3,2) would get size %d" % a)
            a = (a - 3) // 2 + 1              # MaxPool2d(3,2)
        return a
    oh, ow = axis(h), axis(w)
    if oh * ow <= 1:
        raise RuntimeError("Discriminator input too small: InstanceNorm needs more than 1 spatial element")
    return oh, ow


def discriminator_forward(sd: Dict[str, Tensor], x: Tensor, taps: Optional[Dict[str, Tensor]] = None) -> Tensor:
    """Discriminator.forward (src/models.py:92-120): 4 x [conv s2, MaxPool(3,2), InstanceNorm (no affine,
    biased var, eps 1e-5), LeakyReLU(0.2)] with the last LeakyReLU replaced by Sigmoid."""
    discriminator_output_hw(x.shape[2], x.shape[3])
    specs = [("model.0", 2), ("model.4", 1), ("model.8", 1), ("model.12", 1)]
    out = x
    for i, (name, pad) in enumerate(specs):
        out = F.conv2d(out, sd[name + ".weight"], sd[name + ".bias"], stride=2, padding=pad)
        out = F.max_pool2d(out, kernel_size=3, stride=2)
        mean = out.mean(dim=(2, 3), keepdim=True)
        var = out.var(dim=(2, 3), unbiased=False, keepdim=True)
        out = (out - mean) * torch.rsqrt(var + BN_EPS)
        out = F.leaky_relu(out, 0.2) if i < 3 else torch.sigmoid(out)
        if taps is not None:
            taps[name] = out
    return out


# --------------------------------------------------------------------------------------------------
# ReconstructionLoss: src/utils.py:173-241  (closed form in SURVEY Appendix C)
# --------------------------------------------------------------------------------------------------
def _depthwise3(x: Tensor, k3: Tensor) -> Tensor:
    return F.conv2d(x, k3.to(x.dtype).expand(x.shape[1], 1, 3, 3), padding=1, groups=x.shape[1])


_PX = torch.tensor([[-5.0, 0.0, 5.0]] * 3)                       # src/utils.py:180-182
_PY = _PX.t().contiguous()                                         # :184-186
_LAP = torch.full((3, 3), -1.0 / 8.0)
_LAP[1, 1] = 1.0                                                   # :190-192


def edge_weights(hr: Tensor) -> Tensor:
    """high_pass_filter (src/utils.py:200-215): max(|Px*HR|,|Py*HR|) re-normalised over the whole
    batch to mean 1 / std 0.2 (unbiased std, :194-198) and clamped to [0, 2]."""
    e0 = torch.maximum(_depthwise3(hr, _PX).abs(), _depthwise3(hr, _PY).abs())
    e = (e0 - e0.mean()) / e0.std() * 0.2 + 1.0
    return e.clamp(0.0, 2.0)


def reconstruction_loss(hr: Tensor, sr: Tensor) -> Tuple[Tensor, Tensor]:
    """ReconstructionLoss.forward(original_images=hr, target_images=sr) (src/utils.py:228-241)."""
    e = edge_weights(hr)
    edge_loss = ((hr - sr).abs() * e).sum() / e.sum()                       # :232-239
    d = _depthwise3(sr, _LAP)
    tv_loss = F.relu((d.abs() * (1.0 - e)).mean())                          # :217-226
    return edge_loss, tv_loss


def reconstruction_loss_grad(hr: Tensor, sr: Tensor) -> Tensor:
    """d(edge_loss + tv_loss)/d(sr), closed form; L is symmetric so the adjoint of
    the depthwise correlation is the same correlation."""
    e = edge_weights(hr)
    n = sr.numel()
    g = torch.sign(sr - hr) * e / e.sum()
    d = _depthwise3(sr, _LAP)
    m = (d.abs() * (1.0 - e)).mean()
    if m > 0:
        g = g + _depthwise3(torch.sign(d) * (1.0 - e), _LAP) / n
    return g


# --------------------------------------------------------------------------------------------------
# Adam (torch.optim.Adam defaults as used at src/train.py:61-62: betas (0.9, 0.999), eps 1e-8, no decay)
# --------------------------------------------------------------------------------------------------
class AdamState:
    def __init__(self, params: List[Tensor], lr: float, betas=(0.9, 0.999), eps: float = 1e-8):
        self.params, self.lr, self.betas, self.eps = params, lr, betas, eps
        self.m = [torch.zeros_like(p) for p in params]
        self.v = [torch.zeros_like(p) for p in params]
        self.t = 0

    def step(self, grads: List[Optional[Tensor]]) -> None:
        self.t += 1
        b1, b2 = self.betas
        c1, c2 = 1 - b1 ** self.t, 1 - b2 ** self.t
        with torch.no_grad():
            for p, g, m, v in zip(self.params, grads, self.m, self.v):
                if g is None:
                    continue
                m.mul_(b1).add_(g, alpha=1 - b1)
                v.mul_(b2).addcmul_(g, g, value=1 - b2)
                denom = (v.sqrt() / math.sqrt(c2)).add_(self.eps)
                p.addcdiv_(m, denom, value=-self.lr / c1)


# --------------------------------------------------------------------------------------------------
# Train steps: src/train.py:175-203 (train_generator), :206-230 (train_discriminator), :184-192 (GAN term)
# --------------------------------------------------------------------------------------------------
PARAM_SUFFIXES = (".weight", ".bias")


def trainable_keys(sd: Dict[str, Tensor]) -> List[str]:
    return [k for k in sd if k.endswith(PARAM_SUFFIXES)]


def _with_grad(sd: Dict[str, Tensor]) -> Dict[str, Tensor]:
    out = {}
    for k, v in sd.items():
        out[k] = v.detach().requires_grad_(True) if k.endswith(PARAM_SUFFIXES) else v
    return out


def generator_loss_and_grads(g_sd: Dict[str, Tensor], lr_imgs: Tensor, hr_imgs: Tensor,
                             d_sd: Optional[Dict[str, Tensor]] = None, gan_mode: bool = False):
    """One generator objective evaluation in train mode.  Pixel mode is HEAD's train_generator
    (g_loss = com_loss + tv_loss, src/train.py:189-192); ``gan_mode`` adds the commented-out adversarial
    term g_d_loss = mean(tanh(D(hr) - D(sr))) (src/train.py:184-190).  Returns (losses, grads-by-key, sr)."""
    work = _with_grad(g_sd)
    sr = srresnet_forward(work, lr_imgs, training=True, update_running=True)
    com, tv = reconstruction_loss(hr_imgs, sr)
    g_d = torch.zeros(())
    if gan_mode:
        fake = discriminator_forward(d_sd, sr)
        with torch.no_grad():
            real = discriminator_forward(d_sd, hr_imgs)
        g_d = torch.tanh(real - fake).mean()
        loss = com + tv + g_d
    else:
        loss = com + tv
    keys = trainable_keys(work)
    grads = torch.autograd.grad(loss, [work[k] for k in keys])
    for k in g_sd:  # running stats were updated on the working copy's shared buffers
        if not k.endswith(PARAM_SUFFIXES):
            g_sd[k] = work[k]
    return (float(loss.detach()), float(com.detach()), float(tv.detach()), float(g_d.detach())), dict(zip(keys, grads)), sr.detach()


def train_generator_step(g_sd, opt: AdamState, lr_imgs, hr_imgs, d_sd=None, gan_mode=False):
    """train_generator (src/train.py:175-203): returns (g_loss, com_loss, tv_loss, g_d_loss)."""
    losses, grads, _ = generator_loss_and_grads(g_sd, lr_imgs, hr_imgs, d_sd, gan_mode)
    opt.step([grads[k] for k in trainable_keys(g_sd)])
    return losses


def discriminator_loss_and_grads(d_sd, g_sd, hr_imgs, lr_imgs):
    """train_discriminator's objective (src/train.py:209-218): generator in eval mode (running BN stats),
    d_loss = mean(tanh(D(sr) - D(hr)))."""
    work = _with_grad(d_sd)
    with torch.no_grad():  # the reference keeps the graph into G; its G grads are discarded
        sr = srresnet_forward(g_sd, lr_imgs, training=False)
    real = discriminator_forward(work, hr_imgs)
    fake = discriminator_forward(work, sr)
    loss = torch.tanh(fake - real).mean()
    keys = trainable_keys(work)
    grads = torch.autograd.grad(loss, [work[k] for k in keys])
    return float(loss.detach()), dict(zip(keys, grads))


def train_discriminator_step(d_sd, opt: AdamState, g_sd, hr_imgs, lr_imgs):
    loss, grads = discriminator_loss_and_grads(d_sd, g_sd, hr_imgs, lr_imgs)
    opt.step([grads[k] for k in trainable_keys(d_sd)])
    return loss


# --------------------------------------------------------------------------------------------------
# PSNR as used by the reference metric (src/utils.py:141-144: skimage PSNR, data_range=1)
# --------------------------------------------------------------------------------------------------
def psnr(a: Tensor, b: Tensor) -> float:
    mse = ((a.double() - b.double()) ** 2).mean().item()
    return float("inf") if mse == 0 else 10.0 * math.log10(1.0 / mse)


def image_enhancer(x: Tensor, factor: float = 1.0) -> Tensor:
    """ImageEnhancer.forward (src/models.py:36-41): x + factor * depthwise Laplacian(x), clamped to [0, 1]."""
    return torch.clamp(x + factor * _depthwise3(x, _LAP), 0.0, 1.0)


# --------------------------------------------------------------------------------------------------
# Seeded default initialisation with the reference's construction order (nn.Conv2d / nn.BatchNorm2d
# defaults; order = src/models.py:53-78 and :91-114) so weights can be regenerated anywhere from a seed.
# --------------------------------------------------------------------------------------------------
def _conv_init(cout: int, cin: int, k: int) -> Tuple[Tensor, Tensor]:
    w = torch.empty(cout, cin, k, k)
    torch.nn.init.kaiming_uniform_(w, a=math.sqrt(5))
    bound = 1.0 / math.sqrt(cin * k * k)
    b = torch.empty(cout).uniform_(-bound, bound)
    return w, b


def init_srresnet_state(seed: int, in_channels=3, num_features=64, num_residuals=16, upscale_factor=4):
    torch.manual_seed(seed)
    sd: Dict[str, Tensor] = {}
    sd["conv1.weight"], sd["conv1.bias"] = _conv_init(num_features, in_channels, 9)
    for i in range(num_residuals):
        p = f"residual_blocks.{i}"
        for c, bn in (("conv1", "bn1"), ("conv2", "bn2")):
            sd[f"{p}.{c}.weight"], sd[f"{p}.{c}.bias"] = _conv_init(num_features, num_features, 3)
            sd[f"{p}.{bn}.weight"] = torch.ones(num_features)
            sd[f"{p}.{bn}.bias"] = torch.zeros(num_features)
            sd[f"{p}.{bn}.running_mean"] = torch.zeros(num_features)
            sd[f"{p}.{bn}.running_var"] = torch.ones(num_features)
            sd[f"{p}.{bn}.num_batches_tracked"] = torch.zeros((), dtype=torch.long)
    sd["conv2.weight"], sd["conv2.bias"] = _conv_init(num_features, num_features, 3)
    for s in range(int(upscale_factor // 2)):
        sd[f"upsample.{3 * s}.weight"], sd[f"upsample.{3 * s}.bias"] = _conv_init(num_features * 4, num_features, 3)
    sd["conv3.weight"], sd["conv3.bias"] = _conv_init(in_channels, num_features, 9)
    return sd


def init_discriminator_state(seed: int, input_channels=3, num_filters=64):
    torch.manual_seed(seed)
    sd: Dict[str, Tensor] = {}
    sd["model.0.weight"], sd["model.0.bias"] = _conv_init(num_filters, input_channels, 8)
    sd["model.4.weight"], sd["model.4.bias"] = _conv_init(num_filters * 2, num_filters, 4)
    sd["model.8.weight"], sd["model.8.bias"] = _conv_init(num_filters * 4, num_filters * 2, 4)
    sd["model.12.weight"], sd["model.12.bias"] = _conv_init(num_filters * 8, num_filters * 4, 4)
    return sd


# ----------------------------------------------------------------------------------------------------------------
# VGG19 perceptual loss (reference: src/models.py:123-151 VGGFeatureExtractor, src/utils.py:154-166 perceptal_loss)
# ----------------------------------------------------------------------------------------------------------------
# torchvision vgg19 'E' configuration; module index i of vgg19.features <-> position in this walk
VGG19_CFG = [64, 64, "M", 128, 128, "M", 256, 256, 256, 256, "M", 512, 512, 512, 512, "M", 512, 512, 512, 512, "M"]
VGG_LAYER_NAMES = {"3": "conv1_2", "8": "conv2_2", "17": "conv3_3", "26": "conv4_3", "35": "conv5_3"}     # src/models.py:128-134


def init_vgg19_state(seed: int, prefix: str = "vgg19.") -> Dict[str, Tensor]:
    """Feature-extractor tensors of torchvision's ``vgg19(weights=None)`` built under ``torch.manual_seed(seed)``: what the
    reference's VGGFeatureExtractor holds when the ImageNet download is replaced by the default initialisation
    (tests/golden/make_golden.py:make_vgg_fixture).  The initialiser order inside torchvision (default Conv2d / Linear
    init, then kaiming_normal_ over all modules) is what fixes the values, so torchvision itself builds them."""
    import torchvision
    torch.manual_seed(seed)
    feats = torchvision.models.vgg19(weights=None).features
    return {prefix + k: v.detach().clone() for k, v in feats.state_dict().items()}


def vgg_features(sd: Dict[str, Tensor], x: Tensor, layers=("conv3_3", "conv4_3"), prefix: str = "vgg19.") -> Dict[str, Tensor]:
    """VGGFeatureExtractor.forward (src/models.py:140-151): walk vgg19.features, collect the selected (post-ReLU) maps,
    stop once all are collected."""
    feats: Dict[str, Tensor] = {}
    want = [n for n in VGG_LAYER_NAMES.values() if n in layers]
    idx = 0
    for v in VGG19_CFG:
        if v == "M":
            x = F.max_pool2d(x, 2, 2)
            steps = [idx]
            idx += 1
        else:
            x = F.conv2d(x, sd[f"{prefix}{idx}.weight"], sd[f"{prefix}{idx}.bias"], padding=1)
            x = F.relu(x)
            steps = [idx, idx + 1]
            idx += 2
        for s in steps:
            name = VGG_LAYER_NAMES.get(str(s))
            if name is not None and name in layers:
                feats[name] = x
        if len(feats) == len(want):
            break
    return feats


def perceptual_loss(sd: Dict[str, Tensor], sr: Tensor, hr: Tensor, layers=("conv3_3", "conv4_3")) -> Tensor:
    """perceptal_loss (src/utils.py:154-166): sum over the selected layers of mean |features(sr) - features(hr)|."""
    fr = vgg_features(sd, hr, layers)
    ff = vgg_features(sd, sr, layers)
    total = 0
    for k in fr:
        total = total + (ff[k] - fr[k]).abs().mean()
    return total
