"""Developer tool (GPU box): per-layer error report of the CUDA generator against the CPU oracle.
Usage: python tools/check_generator.py [N H W [num_residuals]]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import srgan_b200 as S  # noqa: E402
from oracle import srgan_oracle as O  # noqa: E402
from oracle import bf16_storage_model as B  # noqa: E402


def rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).abs().max() / max(b.abs().max(), 1e-30)), float((a - b).norm() / max(b.norm(), 1e-30))


def nhwc_to_nchw(t):
    return t.float().permute(0, 3, 1, 2).contiguous()


def main():
    args = [int(a) for a in sys.argv[1:]]
    N, H, W = (args + [2, 16, 24])[:3] if len(args) >= 3 else (2, 16, 24)
    n_res = args[3] if len(args) > 3 else 16
    torch.manual_seed(1)
    g = S.SRResNet(num_residuals=n_res)
    sd = {k: v.clone() for k, v in g.state_dict().items()}
    lr = torch.rand(N, 3, H, W)
    hr = torch.rand(N, 3, 4 * H, 4 * W)
    g = g.cuda()

    # ---- eval forward
    g.eval()
    with torch.no_grad():
        y = g(lr.cuda())
    torch.cuda.synchronize()
    with torch.no_grad():
        y_ref = O.srresnet_forward(sd, lr, training=False)
    print("eval forward  max-rel %.3e  l2-rel %.3e   psnr(ours,ref)=%.2f dB" % (*rel(y, y_ref), O.psnr(y.cpu(), y_ref)))

    # ---- train forward + loss + backward
    g.train()
    crit = S.ReconstructionLoss()
    sr = g(lr.cuda())
    com, tv = crit(hr.cuda(), sr)
    (com + tv).backward()
    torch.cuda.synchronize()
    taps = {}
    sd_t = {k: v.clone() for k, v in sd.items()}
    work = O._with_grad(sd_t)
    sr_ref = O.srresnet_forward(work, lr, training=True, update_running=True, taps=taps)
    com_r, tv_r = O.reconstruction_loss(hr, sr_ref)
    keys = O.trainable_keys(work)
    grads_ref = dict(zip(keys, torch.autograd.grad(com_r + tv_r, [work[k] for k in keys])))
    print("train forward max-rel %.3e  l2-rel %.3e" % rel(sr.detach(), sr_ref.detach()))
    print("com %.7f (ref %.7f)   tv %.3e (ref %.3e)" % (float(com), float(com_r), float(tv), float(tv_r)))
    eng = g.last_engine()
    names = {"out1": "out1", "trunk": "trunk"}
    for b in range(n_res):
        names[f"rb{b}.y1"] = f"residual_blocks.{b}.conv1"
        names[f"rb{b}.y2"] = f"residual_blocks.{b}.conv2"
        names[f"rb{b}.out"] = f"residual_blocks.{b}"
    for j in range(g.num_upsample_stages):
        names[f"up{j}"] = f"upsample.{3 * j}"
    for en, on in names.items():
        t = nhwc_to_nchw(eng.named_tensor(en))
        print("  tap %-28s max-rel %.3e  l2-rel %.3e" % (en, *rel(t, taps[on].detach())))
    # bf16-storage model of the same arithmetic (diagnostic)
    work_b = O._with_grad({k: v.clone() for k, v in sd.items()})
    taps_b = {}
    sr_b = B.srresnet_forward_train(work_b, lr, taps=taps_b)
    com_b, tv_b = O.reconstruction_loss(hr, sr_b)
    grads_b = dict(zip(keys, torch.autograd.grad(com_b + tv_b, [work_b[k] for k in keys])))
    print("train forward vs bf16-storage model: max-rel %.3e l2-rel %.3e" % rel(sr.detach(), sr_b.detach()))
    for en, on in names.items():
        t = nhwc_to_nchw(eng.named_tensor(en))
        print("  tapB %-28s max-rel %.3e  l2-rel %.3e" % (en, *rel(t, taps_b[on].detach())))
    worst = worst_b = 0.0
    for k, p in g.named_parameters():
        mr, l2 = rel(p.grad, grads_ref[k])
        mrb, l2b = rel(p.grad, grads_b[k])
        if not k.endswith("conv1.bias") and not k.endswith("conv2.bias") or k in ("conv1.bias", "conv2.bias"):
            worst = max(worst, l2)
            worst_b = max(worst_b, l2b)
        print("  grad %-40s fp32-oracle l2-rel %.3e | bf16-model max-rel %.3e l2-rel %.3e  |ref| %.3e" % (
            k, l2, mrb, l2b, float(grads_ref[k].norm())))
    print("worst grad l2-rel: vs fp32 oracle %.3e, vs bf16-storage model %.3e" % (worst, worst_b))
    for k in ("residual_blocks.0.bn1.running_mean", "residual_blocks.0.bn1.running_var"):
        if k in work:
            print("  %s max-rel %.3e" % (k, rel(g.state_dict()[k], work[k])[0]))

    # ---- loss gradient alone (fp32 path)
    srd = sr_ref.detach().cuda().requires_grad_(True)
    c2, t2 = crit(hr.cuda(), srd)
    (c2 + t2).backward()
    gl_ref = O.reconstruction_loss_grad(hr, sr_ref.detach())
    print("loss value diff %.3e %.3e ; loss grad max-rel %.3e l2-rel %.3e" % (
        abs(float(c2) - float(com_r)), abs(float(t2) - float(tv_r)), *rel(srd.grad, gl_ref)))

    # ---- Adam
    opt = S.Adam(g.parameters(), lr=1e-4)
    before = g.flat_parameters().clone()
    gflat = g.flat_grads().clone()
    opt.step()
    torch.cuda.synchronize()
    m = 0.1 * gflat
    v = 0.001 * gflat * gflat
    exp = before - (1e-4 / 0.1) * m / (v.sqrt() / (0.001 ** 0.5) + 1e-8)
    print("adam step max abs diff %.3e" % float((g.flat_parameters() - exp).abs().max()))
    print("launches so far:", g.launch_count())


if __name__ == "__main__":
    main()
