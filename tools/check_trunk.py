"""Developer tool (GPU box): A/B of the fused trunk kernel (csrc/trunk_fused.cu) against the per-layer launch path.
Same weights, same inputs: every named intermediate, the output and every parameter gradient of one train-mode
forward + backward are compared, the bounded-wait error word is read back, and both paths are timed.
Usage: python tools/check_trunk.py [N H W]..."""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import srgan_b200 as S  # noqa: E402


def rel(a, b):
    a, b = a.double(), b.double()
    return float((a - b).abs().max() / max(float(b.abs().max()), 1e-30))


def prof_dump(tag, K=1):
    import ctypes
    import numpy as np
    NT = 148 * 18 + 148 * 128 * 16
    buf = (ctypes.c_longlong * NT)()
    n = S.lib().srg_debug_trunk_prof(buf, NT)
    if n <= 0:
        return
    full = np.array(buf[:n], dtype=np.int64)
    tot = full[:148 * 18].reshape(148, 18)[:, 4].max()
    tr = full[148 * 18:].reshape(148, 128, 16)
    names = ["prod:flags ok", "mma:operands in", "mma:last issued", "st:raw stored", "p1:slot done", "ap:arrive", "ap:sums published",
             "ap:transform start"]
    print(f"  [{tag}] kernel cycles (max over CTAs) {tot}")
    for cta in (0, 75, 140):
        for s_ in (3 * K, 3 * K + 1, 4 * K):
            t0 = tr[cta, s_, 1]
            print(f"    cta {cta} slot {s_}: " + "  ".join(f"{names[e].split(':')[1]}={tr[cta, s_, e] - t0}" for e in (0, 2, 4, 3, 5, 6, 7))
                  + f"  | next slot: flags ok={tr[cta, s_ + 1, 0] - t0} operands in={tr[cta, s_ + 1, 1] - t0}"
                  + (f"  | same generator's next layer: flags ok={tr[cta, s_ + K, 0] - t0} operands in={tr[cta, s_ + K, 1] - t0}" if K > 1 else ""))
            e = tr[cta, s_, 8:16] - tr[cta, s_, 8]
            print(f"      tile 1 epilogue: pass2 {e[1]}  bar_a+y {e[2] - e[1]}  wait acc {e[3] - e[2]}  wait buf {e[4] - e[3]}  tmem+stage+park {e[5] - e[4]}  bar_b {e[6] - e[5]}  stats {e[7] - e[6]}  total {e[7]}")


def one_pass(g, crit, lr, hr):
    g.zero_grad(set_to_none=True)
    sr = g(lr)
    com, tv = crit(hr, sr)
    (com + tv).backward()
    torch.cuda.synchronize()
    eng = g.last_engine()
    T = {name: eng.named_tensor(name).float().clone() for name in eng.tensor_table()}
    G = {k: p.grad.detach().clone() for k, p in g.named_parameters()}
    return sr.detach().clone(), T, G, float(com), float(tv)


def timeit(g, crit, lr, hr, n=10):
    for _ in range(3):
        sr = g(lr); c, t = crit(hr, sr); (c + t).backward()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        sr = g(lr); c, t = crit(hr, sr); (c + t).backward()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def check(N, H, W, keep):
    L = S.lib()
    torch.manual_seed(1)
    g = S.SRResNet().cuda()
    g.debug_keep_grads = keep
    g.train()
    crit = S.ReconstructionLoss()
    lr = torch.rand(N, 3, H, W, device="cuda")
    hr = torch.rand(N, 3, 4 * H, 4 * W, device="cuda")
    L.srg_set_trunk_fused(0)
    sr0, T0, G0, c0, t0 = one_pass(g, crit, lr, hr)
    L.srg_set_trunk_fused(1)
    sr1, T1, G1, c1, t1 = one_pass(g, crit, lr, hr)
    eng = g.last_engine()
    layers = L.srg_generator_trunk_layers(eng.handle)
    err = L.srg_generator_trunk_error(eng.handle)
    print(f"--- {N}x3x{H}x{W} keep_grads={keep}: fused launch covers {layers} layers, error word {err}")
    print(f"  sr max-rel {rel(sr1, sr0):.3e}   com {c1:.7f} vs {c0:.7f}   tv {t1:.4e} vs {t0:.4e}")
    worst = {}
    for k in T0:
        cls = k.split(".")[-1] if k.startswith("rb") else k
        r = rel(T1[k], T0[k])
        if r > worst.get(cls, (0, ""))[0] or cls not in worst:
            worst[cls] = (r, k)
    for cls, (r, k) in sorted(worst.items()):
        print(f"  tensor class {cls:12s} worst max-rel {r:.3e} ({k})")
    wg = max((rel(G1[k], G0[k]), k) for k in G0 if float(G0[k].abs().max()) > 0)
    print(f"  worst parameter-gradient max-rel {wg[0]:.3e} ({wg[1]})")
    for k in ("residual_blocks.15.bn2.weight", "residual_blocks.0.bn1.bias", "residual_blocks.0.conv1.weight", "conv1.weight"):
        print(f"    grad {k:36s} max-rel {rel(G1[k], G0[k]):.3e}")
    sd = g.state_dict()
    print("  running_mean[rb0.bn1][:4]", sd["residual_blocks.0.bn1.running_mean"][:4].tolist())
    if not keep:
        L.srg_set_trunk_fused(0)
        ms0 = timeit(g, crit, lr, hr)
        L.srg_set_trunk_fused(1)
        ms1 = timeit(g, crit, lr, hr)
        print(f"  eager fwd+loss+bwd: per-layer launches {ms0:.3f} ms, fused trunk {ms1:.3f} ms; error word "
              f"{L.srg_generator_trunk_error(g.last_engine().handle)}")
        for fused in (0, 1):
            L.srg_set_trunk_fused(fused)
            with torch.no_grad():
                for _ in range(3):
                    g(lr)
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(10):
                    g(lr)
                e1.record()
                torch.cuda.synchronize()
            print(f"  train-mode forward only, fused={fused}: {e0.elapsed_time(e1) / 10:.3f} ms")
            if fused:
                prof_dump("fwd")
        g.profile_enable(True)
        sr = g(lr); c, t = crit(hr, sr); (c + t).backward()
        torch.cuda.synchronize()
        ms, n = g.profile_read()
        g.profile_enable(False)
        print(f"  fused trunk launches this step: {n}, {ms:.3f} ms in total (fwd + bwd)")
        prof_dump("bwd")
    return err


def main():
    args = [int(a) for a in sys.argv[1:]]
    geos = [tuple(args[i:i + 3]) for i in range(0, len(args) - 2, 3)] or [(2, 16, 24), (3, 40, 20), (8, 64, 64), (16, 96, 96)]
    bad = 0
    for (N, H, W) in geos:
        t = time.time()
        bad |= check(N, H, W, True)
        bad |= check(N, H, W, False)
        print(f"  ({time.time() - t:.1f} s)", flush=True)
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()
