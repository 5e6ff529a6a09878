"""GPU parity tests of the VGG19 perceptual loss (src/models.py:123-151, src/utils.py:154-166): the helper
kernels one by one against torch on the same bf16 values, then VGGFeatureExtractor / perceptal_loss end to end against the
fixture recorded from the unmodified reference (tests/golden/make_golden.py vgg) and the CPU oracle."""
import os
from ctypes import c_void_p

import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def S():
    import srgan_b200
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return srgan_b200


@pytest.fixture(scope="module")
def O():
    from oracle import srgan_oracle
    return srgan_oracle


def P(t):
    return c_void_p(t.data_ptr())


def maxrel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).abs().max() / max(float(b.abs().max()), 1e-30))


def nhwc(x):
    return x.permute(0, 2, 3, 1).contiguous().to(torch.bfloat16).cuda()


def nchw(t):
    return t.float().permute(0, 3, 1, 2).contiguous().cpu()


def rb(x):
    return x.to(torch.bfloat16).float()


@pytest.mark.parametrize("shape", [(2, 3, 9, 7), (1, 3, 16, 24), (3, 3, 5, 33)])
def test_first_layer_as_unfold_plus_1x1_matches_conv3x3(S, shape):
    """vgg19.features[0] (3 -> 64, 3x3, pad 1) = srg_unfold3x3_rgb + a 1x1 srg_conv2d_fprop; srg_fold3x3_rgb is the adjoint."""
    L = S.lib()
    N, _, H, W = shape
    torch.manual_seed(H * W)
    x = torch.rand(*shape)
    w = torch.randn(64, 3, 3, 3) * 0.2
    b = torch.randn(64) * 0.1
    unf = torch.empty(N, H, W, 64, dtype=torch.bfloat16, device="cuda")
    xc, bc = x.cuda(), b.cuda()
    S.check(L.srg_unfold3x3_rgb(P(xc), N, H, W, P(unf), None), "unfold")
    # the unfolded tensor is exactly F.unfold's patches in (kh, kw, c) channel order, rounded to bf16
    ref_unf = F.unfold(x, 3, padding=1).reshape(N, 3, 9, H, W).permute(0, 3, 4, 2, 1).reshape(N, H, W, 27)
    assert torch.equal(unf[..., :27].float().cpu(), rb(ref_unf))
    assert float(unf[..., 27:].float().abs().max()) == 0.0
    w1 = torch.zeros(64, 64, 1, 1)
    w1[:, :27, 0, 0] = w.permute(0, 2, 3, 1).reshape(64, 27)
    wf = torch.empty(int(L.srg_conv2d_packed_weight_bytes(64, 64, 1)), dtype=torch.uint8, device="cuda")
    w1c = w1.cuda()
    S.check(L.srg_conv2d_pack_weights(P(w1c), 64, 64, 1, 0, P(wf), None), "pack")
    out = torch.empty(N, H, W, 64, dtype=torch.bfloat16, device="cuda")
    S.check(L.srg_conv2d_fprop(P(unf), N, H, W, 64, P(wf), 64, 1, P(bc), 1, 0.0, None, P(out), None), "fprop")
    ref = F.relu(F.conv2d(rb(x), rb(w), b, padding=1))
    assert maxrel(nchw(out), ref) < 1e-2
    # adjoint: <unfold(x), g> == <x, fold(g)> ; compare fold with autograd through F.unfold
    g = torch.randn(N, H, W, 64)
    g[..., 27:] = 0
    gd = g.to(torch.bfloat16).cuda()
    dx = torch.empty(N, 3, H, W, device="cuda")
    S.check(L.srg_fold3x3_rgb(P(gd), N, H, W, 0.5, P(dx), None), "fold")
    xr = x.clone().requires_grad_(True)
    u = F.unfold(xr, 3, padding=1).reshape(N, 3, 9, H, W).permute(0, 3, 4, 2, 1).reshape(N, H, W, 27)
    (u * rb(g[..., :27])).sum().backward()
    assert maxrel(dx, 0.5 * xr.grad) < 1e-5


@pytest.mark.parametrize("shape", [(2, 64, 8, 12), (1, 128, 9, 7), (2, 256, 5, 6)])
def test_maxpool2x2_forward_backward_bit_exact(S, shape):
    """nn.MaxPool2d(2, 2) on bf16 values is exact: forward and the routed gradient (first maximum) equal torch's bit for
    bit, including odd sizes (floor) and the optional second gradient."""
    L = S.lib()
    N, C, H, W = shape
    torch.manual_seed(C + H)
    x = rb(torch.randn(*shape))
    x[:, :, : H // 2 * 2 : 2, : W // 2 * 2 : 2] = x[:, :, 1 : H // 2 * 2 : 2, 1 : W // 2 * 2 : 2]     # ties inside every window
    xd = nhwc(x)
    out = torch.empty(N, H // 2, W // 2, C, dtype=torch.bfloat16, device="cuda")
    S.check(L.srg_maxpool2x2_forward(P(xd), N, H, W, C, P(out), None), "pool fwd")
    xr = x.clone().requires_grad_(True)
    ref = F.max_pool2d(xr, 2, 2)
    assert torch.equal(nchw(out), ref.detach())
    dy = rb(torch.randn_like(ref))
    add = rb(torch.randn_like(x))
    ref.backward(dy)
    dx = torch.empty(N, H, W, C, dtype=torch.bfloat16, device="cuda")
    dyd, addd = nhwc(dy), nhwc(add)                  # keep the device tensors alive across the raw-pointer calls
    S.check(L.srg_maxpool2x2_backward(P(xd), P(dyd), None, N, H, W, C, P(dx), None), "pool bwd")
    assert torch.equal(nchw(dx), xr.grad)
    S.check(L.srg_maxpool2x2_backward(P(xd), P(dyd), P(addd), N, H, W, C, P(dx), None), "pool bwd")
    assert torch.equal(nchw(dx), rb(xr.grad + add))
    assert L.srg_maxpool2x2_forward(P(xd), N, 1, W, C, P(out), None) != 0          # smaller than the window: error code


def test_l1_bf16_value_and_masked_gradient(S):
    L = S.lib()
    torch.manual_seed(3)
    n = 8 * 5000
    a = rb(torch.relu(torch.randn(n)))
    b = rb(torch.relu(torch.randn(n)))
    b[:100] = a[:100]                                           # exact ties: zero gradient
    ad, bd = a.to(torch.bfloat16).cuda(), b.to(torch.bfloat16).cuda()
    scratch = torch.empty(int(L.srg_l1_bf16_scratch_bytes()), dtype=torch.uint8, device="cuda")
    out = torch.full((1,), 0.25, device="cuda")
    g = torch.empty(n, dtype=torch.bfloat16, device="cuda")
    S.check(L.srg_l1_bf16(P(ad), P(bd), n, 2.0, 1, 3.0, 1, P(g), P(scratch), scratch.numel(), P(out), None), "l1")
    ar = a.clone().requires_grad_(True)
    ref = F.l1_loss(torch.relu(ar), b)          # a is a ReLU output: the mask is the ReLU's derivative
    ref.backward()
    assert abs(float(out) - (0.25 + 2.0 * float(ref))) < 1e-6
    assert maxrel(g.float(), rb(6.0 * ar.grad)) < 1e-6
    S.check(L.srg_l1_bf16(P(ad), P(bd), n, 1.0, 0, 1.0, 0, None, P(scratch), scratch.numel(), P(out), None), "l1")
    assert abs(float(out) - float(ref)) < 1e-6
    assert L.srg_l1_bf16(P(ad), P(bd), 12, 1.0, 0, 1.0, 0, None, P(scratch), scratch.numel(), P(out), None) != 0


def test_feature_extractor_and_perceptual_loss_match_reference_golden(S, O, golden_dir):
    z = np.load(os.path.join(golden_dir, "vgg_perceptual.npz"))
    sd = O.init_vgg19_state(int(z["seed"]))
    fe = S.VGGFeatureExtractor()
    assert sorted(fe.state_dict().keys()) == list(z["keys"])                     # the reference's state_dict keys
    assert all(not p.requires_grad for p in fe.parameters())
    fe.load_state_dict({k: v for k, v in sd.items()}, strict=False)
    missing = [k for k in fe.state_dict() if k not in sd]
    assert all(int(k.split(".")[1]) > 26 for k in missing) or not missing
    fe = fe.cuda()
    sr = torch.from_numpy(z["sr"]).cuda().requires_grad_(True)
    hr = torch.from_numpy(z["hr"]).cuda()
    feats = fe(hr)
    assert set(feats) == {"conv3_3", "conv4_3"}
    # 8 / 12 stacked bf16 conv layers end to end against the fp32 reference
    assert maxrel(feats["conv3_3"], torch.from_numpy(z["conv3_3"])) < 3e-2
    assert maxrel(feats["conv4_3"], torch.from_numpy(z["conv4_3"])) < 3e-2
    loss = S.perceptal_loss(sr, hr, fe)
    loss.backward()
    assert abs(float(loss) - float(z["loss"])) < 1e-2 * float(z["loss"])
    g, g_ref = sr.grad.cpu().double().flatten(), torch.from_numpy(z["grad"]).double().flatten()
    cos = float((g * g_ref).sum() / (g.norm() * g_ref.norm()))
    assert cos > 0.97, cos
    assert 0.9 < float(g.norm() / g_ref.norm()) < 1.1
    # the CPU oracle on the same inputs (what the fixture pins) agrees with the fixture
    assert abs(float(O.perceptual_loss(sd, torch.from_numpy(z["sr"]), torch.from_numpy(z["hr"]))) - float(z["loss"])) < 1e-6
    # upstream gradient scaling and CPU inputs
    sr2 = torch.from_numpy(z["sr"]).cuda().requires_grad_(True)
    (2.5 * S.perceptal_loss(sr2, hr, fe)).backward()
    assert maxrel(sr2.grad, 2.5 * sr.grad) < 1e-6
    with pytest.raises(RuntimeError):
        fe(hr.cpu())


def test_single_layer_selection_and_early_exit(S, O):
    """layers=('conv2_2',): the walk stops after index 8 (src/models.py:149-150); one pooled stage, 4 convs."""
    sd = O.init_vgg19_state(5)
    fe = S.VGGFeatureExtractor(layers=("conv2_2",))
    fe.load_state_dict(sd)
    fe = fe.cuda()
    torch.manual_seed(8)
    x = torch.rand(1, 3, 24, 16)
    ref = O.vgg_features(sd, x, layers=("conv2_2",))["conv2_2"]
    n0 = S.lib().srg_total_launches()
    out = fe(x.cuda())
    assert list(out) == ["conv2_2"] and out["conv2_2"].shape == ref.shape
    assert S.lib().srg_total_launches() - n0 <= 1 + 4 + 1 + 8                      # unfold + 4 convs + pool (+ weight packs)
    assert maxrel(out["conv2_2"], ref) < 2e-2
