"""GPU parity tests of the per-operator convolution C-ABI (include/srgan_b200.h: srg_conv2d_*) against
torch.nn.functional.conv2d in fp32 on the SAME bf16-rounded operands (north_star: per-layer max relative error <= 1e-2,
relative to the largest magnitude of the reference tensor; weight gradients from identical inputs 5e-3)."""
from ctypes import c_void_p

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


def maxrel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).abs().max() / max(float(b.abs().max()), 1e-30))


def P(t):
    return c_void_p(t.data_ptr())


def nhwc_bf16(x):       # NCHW fp32 (cpu) -> NHWC bf16 (cuda)
    return x.permute(0, 2, 3, 1).contiguous().to(torch.bfloat16).cuda()


def nchw_f32(t):        # NHWC bf16 (cuda) -> NCHW fp32 (cpu)
    return t.float().permute(0, 3, 1, 2).contiguous().cpu()


def rb(x):              # the value the kernels see
    return x.to(torch.bfloat16).float()


def pack(S, w, dgrad):
    L = S.lib()
    cout, cin, k, _ = w.shape
    buf = torch.empty(int(L.srg_conv2d_packed_weight_bytes(cout, cin, k)), dtype=torch.uint8, device="cuda")
    wd = w.cuda().contiguous()
    S.check(L.srg_conv2d_pack_weights(P(wd), cout, cin, k, 1 if dgrad else 0, P(buf), None), "pack")
    torch.cuda.synchronize()
    return buf


CASES = [
    # N, H, W, cin, cout, k     (VGG19-like channel counts; ragged tiles; conv3_il and generic / streamed-weight kernels)
    (2, 24, 20, 64, 64, 3),
    (1, 33, 17, 64, 128, 3),
    (2, 16, 24, 128, 256, 3),
    (1, 20, 12, 256, 512, 3),
    (1, 12, 20, 512, 512, 3),
    (2, 16, 16, 64, 64, 1),
    (1, 18, 10, 192, 128, 1),
]


@pytest.mark.parametrize("N,H,W,cin,cout,k", CASES)
def test_conv2d_fprop_and_dgrad_match_torch(S, N, H, W, cin, cout, k):
    L = S.lib()
    torch.manual_seed(N * 1000 + cin + cout + k)
    x = torch.randn(N, cin, H, W)
    w = torch.randn(cout, cin, k, k) / (cin * k * k) ** 0.5
    b = torch.randn(cout)
    res = torch.randn(N, cout, H, W)
    xd, rd, bd = nhwc_bf16(x), nhwc_bf16(res), b.cuda()
    out = torch.empty(N, H, W, cout, dtype=torch.bfloat16, device="cuda")
    wf = pack(S, w, False)
    # forward: ReLU(conv + bias), and conv + bias + residual
    S.check(L.srg_conv2d_fprop(P(xd), N, H, W, cin, P(wf), cout, k, P(bd), 1, 0.0, None, P(out), None), "fprop")
    ref = F.relu(F.conv2d(rb(x), rb(w), b, padding=k // 2))
    assert maxrel(nchw_f32(out), ref) < 1e-2
    S.check(L.srg_conv2d_fprop(P(xd), N, H, W, cin, P(wf), cout, k, P(bd), 0, 0.0, P(rd), P(out), None), "fprop")
    ref = F.conv2d(rb(x), rb(w), b, padding=k // 2) + rb(res)
    assert maxrel(nchw_f32(out), ref) < 1e-2
    # LeakyReLU(0.2)
    S.check(L.srg_conv2d_fprop(P(xd), N, H, W, cin, P(wf), cout, k, None, 2, 0.2, None, P(out), None), "fprop")
    ref = F.leaky_relu(F.conv2d(rb(x), rb(w), None, padding=k // 2), 0.2)
    assert maxrel(nchw_f32(out), ref) < 1e-2
    # input gradient, with the ReLU mask of the layer that produced x
    dy = torch.randn(N, cout, H, W)
    dyd = nhwc_bf16(dy)
    wdg = pack(S, w, True)
    dx = torch.empty(N, H, W, cin, dtype=torch.bfloat16, device="cuda")
    S.check(L.srg_conv2d_dgrad(P(dyd), N, H, W, cout, P(wdg), cin, k, None, None, P(dx), None), "dgrad")
    xr = rb(x).requires_grad_(True)
    F.conv2d(xr, rb(w), None, padding=k // 2).backward(rb(dy))
    assert maxrel(nchw_f32(dx), xr.grad) < 1e-2
    S.check(L.srg_conv2d_dgrad(P(dyd), N, H, W, cout, P(wdg), cin, k, P(xd), None, P(dx), None), "dgrad")
    assert maxrel(nchw_f32(dx), xr.grad * (rb(x) > 0)) < 1e-2


@pytest.mark.parametrize("N,H,W,cin,cout", [(2, 24, 20, 64, 64), (1, 33, 17, 128, 64), (2, 16, 24, 64, 320), (1, 20, 12, 256, 512)])
def test_conv2d_wgrad_matches_torch(S, N, H, W, cin, cout):
    L = S.lib()
    torch.manual_seed(cin * 7 + cout)
    x = torch.randn(N, cin, H, W)
    dy = torch.randn(N, cout, H, W)
    xd, dyd = nhwc_bf16(x), nhwc_bf16(dy)
    nbytes = int(L.srg_conv2d_wgrad_workspace_bytes(N, H, W, cin, cout))
    ws = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    dw = torch.full((cout, cin, 3, 3), float("nan"), device="cuda")
    db = torch.full((cout,), float("nan"), device="cuda")
    S.check(L.srg_conv2d_wgrad(P(xd), P(dyd), N, H, W, cin, cout, P(ws), nbytes, P(dw), P(db), None), "wgrad")
    w = torch.zeros(cout, cin, 3, 3, requires_grad=True)
    b = torch.zeros(cout, requires_grad=True)
    F.conv2d(rb(x), w, b, padding=1).backward(rb(dy))
    assert maxrel(dw, w.grad) < 5e-3
    assert maxrel(db, b.grad) < 1e-4
    # too small a workspace is an error code, not a crash
    assert L.srg_conv2d_wgrad(P(xd), P(dyd), N, H, W, cin, cout, P(ws), 16, P(dw), None, None) != 0


def test_conv2d_argument_checks(S):
    L = S.lib()
    t = torch.zeros(1, 8, 8, 64, dtype=torch.bfloat16, device="cuda")
    assert L.srg_conv2d_fprop(P(t), 1, 8, 8, 48, P(t), 64, 3, None, 0, 0.0, None, P(t), None) != 0      # cin % 64
    assert L.srg_conv2d_fprop(P(t), 1, 8, 8, 64, P(t), 64, 5, None, 0, 0.0, None, P(t), None) != 0      # kernel size
    assert L.srg_conv2d_fprop(P(t), 0, 8, 8, 64, P(t), 64, 3, None, 0, 0.0, None, P(t), None) != 0      # empty
    assert b"multiples of 64" in L.srg_last_error() or b"empty" in L.srg_last_error()
