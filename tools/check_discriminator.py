"""Developer tool (GPU box): per-stage error report of the CUDA discriminator against the CPU oracle.
Usage: python tools/check_discriminator.py [N H W]"""
import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import srgan_b200 as S  # noqa: E402
from oracle import srgan_oracle as O  # noqa: E402


def rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).abs().max() / max(float(b.abs().max()), 1e-30)), float((a - b).norm() / max(float(b.norm()), 1e-30))


def nchw(t):
    return t.float().permute(0, 3, 1, 2).contiguous().cpu()


def uns2d(xs, Hp, Wp, C):
    """XS[n][h'][w'][(a,b,c)] = z[2h'+a-1][2w'+b-1][c]  ->  z (NCHW); positions the conv never reads come back 0."""
    n, Hs, Ws, _ = xs.shape
    xs = xs.float().cpu().view(n, Hs, Ws, 2, 2, C)
    z = torch.zeros(n, C, Hp, Wp)
    for a in range(2):
        for b in range(2):
            hs = [h for h in range(Hs) if 0 <= 2 * h + a - 1 < Hp]
            ws = [w for w in range(Ws) if 0 <= 2 * w + b - 1 < Wp]
            if hs and ws:
                blk = xs[:, hs[0]:hs[-1] + 1, ws[0]:ws[-1] + 1, a, b, :].permute(0, 3, 1, 2)
                z[:, :, 2 * hs[0] + a - 1:2 * hs[-1] + a:2, 2 * ws[0] + b - 1:2 * ws[-1] + b:2] = blk
    return z


def stage_fn(z, w, b, l):
    y = F.conv2d(z, w, b, stride=2, padding=2 if l == 0 else 1)
    p = F.max_pool2d(y, 3, 2)
    xh = F.instance_norm(p, eps=1e-5)
    return y, p, (F.leaky_relu(xh, 0.2) if l < 3 else torch.sigmoid(xh))


def main():
    a = [int(v) for v in sys.argv[1:]]
    N, H, W = a[:3] if len(a) >= 3 else (2, 428, 684)
    torch.manual_seed(3)
    d = S.Discriminator()
    sd = {k: v.clone() for k, v in d.state_dict().items()}
    x = torch.rand(N, 3, H, W)
    d = d.cuda()
    xg = x.cuda().requires_grad_(True)
    out = d(xg)
    torch.manual_seed(5)
    dout = torch.randn(out.shape) * 1e-2
    out.backward(dout.cuda())
    torch.cuda.synchronize()
    eng = d.last_engine()
    T = {k: eng.named_tensor(k) for k in eng.tensor_table()}
    with torch.no_grad():
        ref = O.discriminator_forward(sd, x)
    print("forward end-to-end max-rel %.3e l2-rel %.3e  out shape %s" % (*rel(out.detach(), ref), tuple(out.shape)))
    keys = ["model.0", "model.4", "model.8", "model.12"]
    C = [3, 64, 128, 256, 512]
    # stage inputs as the engine saw them
    zin = [x]
    for l in range(1, 4):
        hp, wp = T[f"p{l - 1}"].shape[1:3]
        zin.append(uns2d(T[f"x{l}"], hp, wp, C[l]))
    for l in range(4):
        w, b = sd[keys[l] + ".weight"], sd[keys[l] + ".bias"]
        y, p, z = stage_fn(zin[l], w, b, l)
        print("stage %d conv  max-rel %.3e l2-rel %.3e" % (l, *rel(nchw(T[f"y{l}"]), y)))
        p_e = F.max_pool2d(nchw(T[f"y{l}"]), 3, 2)
        print("stage %d pool  max-rel %.3e (vs pool of engine conv)" % (l, rel(nchw(T[f"p{l}"]), p_e)[0]))
        xh = F.instance_norm(nchw(T[f"p{l}"]), eps=1e-5)
        z_e = F.leaky_relu(xh, 0.2) if l < 3 else torch.sigmoid(xh)
        if l < 3:
            got = zin[l + 1]
            mask = (got != 0).float()   # rows/cols the next conv never reads are not materialised
            print("stage %d act   max-rel %.3e (vs IN+LReLU of engine pool)" % (l, rel(got, z_e * mask)[0]))
        else:
            print("stage %d out   max-rel %.3e (vs IN+sigmoid of engine pool)" % (l, rel(out.detach(), z_e)[0]))
    # ---- backward, stage by stage, in two isolated halves (same inputs on both sides):
    #   tail: y (engine fp32 conv output) -> pool -> InstanceNorm -> act, upstream = engine's dz
    #   conv: (z_in, W, b) with upstream = engine's dY
    dz = dout
    pw = dict(d.named_parameters())
    for l in range(3, -1, -1):
        y = nchw(T[f"y{l}"]).clone().requires_grad_(True)
        xh = F.instance_norm(F.max_pool2d(y, 3, 2), eps=1e-5)
        z = F.leaky_relu(xh, 0.2) if l < 3 else torch.sigmoid(xh)
        if l < 3:
            z = z * (zin[l + 1] != 0).float()
        z.backward(dz)
        print("stage %d dY    max-rel %.3e l2-rel %.3e   (pool/IN/act backward, engine y)" % (l, *rel(nchw(T[f"dy{l}"]), y.grad)))
        w = sd[keys[l] + ".weight"].clone().requires_grad_(True)
        b = sd[keys[l] + ".bias"].clone().requires_grad_(True)
        zi = zin[l].clone().requires_grad_(True)
        yy = F.conv2d(zi, w, b, stride=2, padding=2 if l == 0 else 1)
        yy.backward(nchw(T[f"dy{l}"]))
        print("stage %d dW    max-rel %.3e l2-rel %.3e ; db max-rel %.3e (|db| %.2e vs |dW| %.2e)" % (
            l, *rel(pw[keys[l] + ".weight"].grad, w.grad), rel(pw[keys[l] + ".bias"].grad, b.grad)[0],
            float(b.grad.abs().max()), float(w.grad.abs().max())))
        if l > 0:
            hp, wp = T[f"p{l - 1}"].shape[1:3]
            dzi = uns2d(T[f"dx{l}"], hp, wp, C[l])
            m2 = (zin[l] != 0).float()
            print("stage %d dZin  max-rel %.3e l2-rel %.3e" % (l, *rel(dzi * m2, zi.grad * m2)))
            dz = dzi
        else:
            print("stage 0 dX    max-rel %.3e l2-rel %.3e" % rel(xg.grad, zi.grad))
    # ---- end-to-end gradient vs oracle autograd (same dout)
    work = O._with_grad(sd)
    xr = x.clone().requires_grad_(True)
    o2 = O.discriminator_forward(work, xr)
    ks = O.trainable_keys(work)
    gr = torch.autograd.grad(o2, [work[k] for k in ks] + [xr], grad_outputs=dout)
    for k, g in zip(ks, gr[:-1]):
        print("e2e grad %-16s max-rel %.3e l2-rel %.3e" % (k, *rel(dict(d.named_parameters())[k].grad, g)))
    print("e2e grad input            max-rel %.3e l2-rel %.3e" % rel(xg.grad, gr[-1]))


if __name__ == "__main__":
    main()
