"""Developer tool (GPU box): the joint multi-generator step (trunks of all K generators in one interleaved launch per
direction, train.joint_pixel_generator_steps) against K independent per-generator graph branches: same initial weights,
same batches -> losses and updated parameters must agree to bf16 noise; both graphs are timed.
Usage: python tools/check_joint.py [K N H W] [--grouped] [--joint-only]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import srgan_b200 as S  # noqa: E402
from tools.check_trunk import prof_dump  # noqa: E402


def make(K, seed0=0):
    gens, opts = [], []
    for s in range(K):
        torch.manual_seed(seed0 + s)
        g = S.SRResNet().cuda()
        g.flat_parameters()
        gens.append(g)
        opts.append(S.Adam(g.parameters(), lr=1e-4, capturable=True))
    return gens, opts


def main():
    joint_only = "--joint-only" in sys.argv
    a = [int(x) for x in sys.argv[1:] if not x.startswith("--")]
    K, N, H, W = (a + [3, 16, 96, 96])[:4] if len(a) >= 4 else (3, 16, 96, 96)
    grouped = "--grouped" in sys.argv       # joint step on the grouped per-layer path (one conv3_il launch per layer for all K)
    # otherwise force the fused trunk kernel (automatic mode picks it for small geometries only)
    S.lib().srg_set_trunk_fused(0 if grouped else 1)
    crit = S.ReconstructionLoss()
    gen = torch.Generator(device="cpu").manual_seed(7)
    batches = [(torch.rand(N, 3, H, W, generator=gen).cuda(), torch.rand(N, 3, 4 * H, 4 * W, generator=gen).cuda()) for _ in range(3)]
    out = {}
    for joint in ((True,) if joint_only else (False, True)):
        gens, opts = make(K)
        step = S.GraphedMultiGeneratorStep(gens, crit, opts, batches[0][0], batches[0][1], joint=joint)
        assert step.joint == joint, "joint path was refused"
        losses = []
        for lr, hr in batches:
            losses.append(step(lr, hr).clone())
        torch.cuda.synchronize()
        flats = [g.flat_parameters().clone() for g in gens]
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for _ in range(3):
            step(*batches[0])
        torch.cuda.synchronize()
        e0.record()
        for _ in range(20):
            step(*batches[0])
        e1.record()
        torch.cuda.synchronize()
        errs = [S.lib().srg_generator_trunk_error(g.last_engine().handle) for g in gens]
        out[joint] = (losses, flats, e0.elapsed_time(e1) / 20, step.launches_per_replay, errs)
        print(f"joint={joint}: {out[joint][2]:.3f} ms per {K}-generator step, {out[joint][3]} launches per replay, error words {errs}",
              flush=True)
        if joint and not grouped:
            prof_dump("joint", K)
    if joint_only:
        sys.exit(1 if any(out[True][4]) else 0)
    la, lb = out[False][0], out[True][0]
    for t in range(3):
        print(f"step {t}: losses per-generator-branches {la[t][:, 0].tolist()}  joint {lb[t][:, 0].tolist()}")
    for i, (fa, fb) in enumerate(zip(out[False][1], out[True][1])):
        d = (fa - fb).abs().max().item()
        print(f"generator {i}: max |param difference| after 3 steps {d:.3e} (max |param| {fa.abs().max().item():.3f}, lr 1e-4)")
    bad = any(out[True][4])
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()
