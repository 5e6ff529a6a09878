"""Developer tool: kernel TIMELINE of the graph-replayed multi-generator step (what ncu cannot show: ncu serialises
launches).  The step is run under torch.profiler (CUPTI activity records: start / end of every kernel of the replayed
CUDA graph, per stream), and summarised as: wall time of the step, per-kernel-name warm in-graph durations, the
fraction of the step during which at least one tensor-core kernel was running, how much elementwise time was hidden
under tensor kernels, and the idle gaps.  Usage: python tools/trace_step.py [--workload cfg2|gan-native] [--out x.json]"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

TENSOR = ("conv3_il", "conv_gemm", "conv9_rows", "wgrad3", "wgrad_gemm", "trunk_kernel", "up_dgrad", "conv9_bwd")


def union_len(iv):
    iv = sorted(iv)
    tot, cur_s, cur_e = 0.0, None, None
    for s, e in iv:
        if cur_e is None or s > cur_e:
            if cur_e is not None:
                tot += cur_e - cur_s
            cur_s, cur_e = s, e
        else:
            cur_e = max(cur_e, e)
    if cur_e is not None:
        tot += cur_e - cur_s
    return tot


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="cfg2")
    ap.add_argument("--batch", type=int, default=16)
    ap.add_argument("--lr-size", type=int, default=96)
    ap.add_argument("--generators", type=int, default=3)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--out", default="gpurun_out/trace_step.json")
    ap.add_argument("--dump", default=None, help="write every kernel record (name, stream, start us, dur us) of one step")
    a = ap.parse_args()
    import torch
    from torch.profiler import ProfilerActivity, profile
    import srgan_b200 as S

    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    K = a.generators
    if a.workload == "gan-native":
        B, LH, LW = 12, 128, 256
    else:
        B, LH, LW = a.batch, a.lr_size, a.lr_size
    gens, opts = [], []
    for s in range(K):
        torch.manual_seed(s)
        g = S.SRResNet().to(dev)
        g.flat_parameters()
        gens.append(g)
        opts.append(S.Adam(g.parameters(), lr=1e-4, capturable=True))
    crit = S.ReconstructionLoss()
    disc, d_opt = None, None
    if a.workload == "gan-native":
        torch.manual_seed(100)
        disc = S.Discriminator().to(dev)
        disc.flat_parameters()
        d_opt = S.Adam(disc.parameters(), lr=5e-5, capturable=True)
    policy = S.MultiGeneratorPolicy(S.PolicyConfig(num_generators=K, force=S.GAN if disc is not None else S.PIXEL, seed=0))
    trainer = S.MultiGeneratorGAN(gens, opts, crit, discriminator=disc, d_optimizer=d_opt, policy=policy, use_cuda_graphs=True)
    lr = torch.rand(B, 3, LH, LW, device=dev)
    hr = torch.rand(B, 3, 4 * LH, 4 * LW, device=dev)
    for _ in range(5):
        trainer.step(lr, hr)
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        for _ in range(a.steps):
            trainer.step(lr, hr)
            torch.cuda.synchronize()
    raw = a.out + ".chrome.json"
    os.makedirs(os.path.dirname(a.out) or ".", exist_ok=True)
    prof.export_chrome_trace(raw)
    with open(raw) as f:
        tr = json.load(f)
    os.remove(raw)
    recs = []
    for e in tr.get("traceEvents", []):
        if e.get("cat") == "kernel" and "dur" in e:
            recs.append((float(e["ts"]), float(e["ts"]) + float(e["dur"]), e["name"], e.get("args", {}).get("stream", 0)))
    recs.sort()
    if not recs:
        print("no CUDA kernel records (CUPTI unavailable?)")
        return
    # split into steps: gaps larger than 200 us separate the synchronised steps
    steps, cur = [], [recs[0]]
    for r in recs[1:]:
        if r[0] - max(x[1] for x in cur[-50:]) > 200.0:
            steps.append(cur)
            cur = [r]
        else:
            cur.append(r)
    steps.append(cur)
    steps = [s for s in steps if len(s) > 50]
    st = steps[-1]
    t0, t1 = min(r[0] for r in st), max(r[1] for r in st)
    wall = t1 - t0

    def short(nm):
        nm = nm.replace("void ", "").replace("srg::", "").replace("(anonymous namespace)::", "").replace("<unnamed>::", "")
        return nm.split("(")[0][:60]

    by = {}
    for s, e, nm, _ in st:
        d = by.setdefault(short(nm), [0, 0.0])
        d[0] += 1
        d[1] += e - s
    tens = [(s, e) for s, e, nm, _ in st if any(t in nm for t in TENSOR)]
    other = [(s, e) for s, e, nm, _ in st if not any(t in nm for t in TENSOR)]
    u_t, u_o, u_all = union_len(tens), union_len(other), union_len(tens + other)
    out = {
        "workload": a.workload, "kernels_in_step": len(st), "wall_us": wall, "steps_seen": len(steps),
        "sum_kernel_us": sum(e - s for s, e, _, _ in st),
        "tensor_kernels_union_us": u_t, "tensor_kernels_sum_us": sum(e - s for s, e in tens),
        "other_kernels_union_us": u_o, "other_kernels_sum_us": sum(e - s for s, e in other),
        "any_kernel_union_us": u_all, "idle_us": wall - u_all,
        "other_not_hidden_under_tensor_us": u_all - u_t,
        "by_kernel": {k: {"n": v[0], "sum_us": round(v[1], 1), "avg_us": round(v[1] / v[0], 2)}
                      for k, v in sorted(by.items(), key=lambda kv: -kv[1][1])},
    }
    os.makedirs(os.path.dirname(a.out) or ".", exist_ok=True)
    with open(a.out, "w") as f:
        json.dump(out, f, indent=1)
    print(json.dumps({k: v for k, v in out.items() if k != "by_kernel"}))
    for k, v in list(out["by_kernel"].items())[:30]:
        print(f"  {k:62s} n={v['n']:4d} sum={v['sum_us']:9.1f} avg={v['avg_us']:8.2f}")
    if a.dump:
        with open(a.dump, "w") as f:
            for s, e, nm, sid in st:
                f.write(f"{s - t0:.2f}\t{e - s:.2f}\t{sid}\t{short(nm)}\n")


if __name__ == "__main__":
    main()
