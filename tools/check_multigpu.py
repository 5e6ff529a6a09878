"""Developer tool (GPU box, torchrun with 2+ ranks): data-parallel step == single-process step on the global batch.

Every rank trains on its shard with SyncBatchNorm + flat-gradient all-reduce (mean).  Rank 0 also runs ONE model on the
concatenated global batch and back-propagates the mean of the per-shard ReconstructionLosses (the
single-process oracle of the sharded step) and compares gradients, losses and BatchNorm running statistics.
Usage: torchrun --nproc-per-node 2 tools/check_multigpu.py"""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import srgan_b200 as S  # noqa: E402


def rel_l2(a, b):
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def check_adversarial(rank, world, dev):
    """Discriminator update and GAN-mode generator step under data parallelism (src/train.py:47 wraps D in DDP,
    :206-230, :184-192) against ONE process on the concatenated global batch: InstanceNorm is per sample, so the rank-mean
    of the per-shard tanh losses / gradients IS the global-batch loss / gradient; the generator runs SyncBatchNorm."""
    b, H, W = 1, 107, 171                              # HR 428 x 684: the smallest valid discriminator input
    torch.manual_seed(1)
    lr_full = torch.rand(world * b, 3, H, W)
    hr_full = torch.rand(world * b, 3, 4 * H, 4 * W)
    torch.manual_seed(20 + rank)
    g = S.SRResNet(num_residuals=2).to(dev)
    torch.manual_seed(30 + rank)
    d = S.Discriminator().to(dev)
    S.parallel.data_parallel([g, d], sync_batchnorm=True, sync_bn_transport="peer")
    lr = S.parallel.shard_batch(lr_full, rank, world).to(dev)
    hr = S.parallel.shard_batch(hr_full, rank, world).to(dev)
    # --- discriminator step (gradients before the optimiser moves anything: lr = 0)
    d_opt = S.Adam(d.parameters(), lr=0.0)
    d_loss = S.train_discriminator_async(d, g, hr, lr, d_opt).clone()
    dist.all_reduce(d_loss); d_loss /= world
    d_flat = d.flat_grads().clone()
    # --- GAN-mode generator step
    g_opt = S.Adam(g.parameters(), lr=0.0)
    crit = S.ReconstructionLoss()
    g_losses = S.train_generator_async(g, d, lr, hr, None, crit, g_opt, gan_mode=True).clone()
    dist.all_reduce(g_losses); g_losses /= world
    g_flat = g.flat_grads().clone()
    torch.cuda.synchronize()
    ok = True
    if rank == 0:
        torch.manual_seed(20); g1 = S.SRResNet(num_residuals=2).to(dev)
        torch.manual_seed(30); d1 = S.Discriminator().to(dev)
        d1_opt = S.Adam(d1.parameters(), lr=0.0)
        l1 = S.train_discriminator_async(d1, g1, hr_full.to(dev), lr_full.to(dev), d1_opt)
        e_dl = abs(float(l1) - float(d_loss))
        e_dg = rel_l2(d_flat, d1.flat_grads())
        g1.train(); d1.eval()
        sr1 = g1(lr_full.to(dev))
        with d1.input_grad_only():
            fake = d1(sr1)
        with torch.no_grad():
            real = d1(hr_full.to(dev))
        loss = S.tanh_mean(real, fake)
        for r in range(world):
            c, t = crit(hr_full[r * b:(r + 1) * b].to(dev), sr1[r * b:(r + 1) * b])
            loss = loss + (c + t) / world
        g1.zero_grad()
        loss.backward()
        torch.cuda.synchronize()
        e_gl = abs(float(loss.detach()) - float(g_losses[0]))
        e_gg = rel_l2(g_flat, g1.flat_grads())
        print(f"adversarial, world={world}: D loss diff {e_dl:.3e}, D grads l2-rel {e_dg:.3e}; GAN-mode G loss diff {e_gl:.3e}, "
              f"G grads l2-rel {e_gg:.3e}; peer_sync_errors={S.parallel.peer_sync_errors()}")
        ok = e_dl < 1e-5 and e_dg < 1e-4 and e_gl < 1e-4 and e_gg < 8e-2
    # --- the same steps replayed from CUDA graphs (captured collectives) == eager, per rank
    def make(graphs):
        torch.manual_seed(40); ga = S.SRResNet(num_residuals=2).to(dev)
        torch.manual_seed(41); da = S.Discriminator().to(dev)
        S.parallel.data_parallel([ga, da], sync_batchnorm=True, sync_bn_transport="peer")
        go = S.Adam(ga.parameters(), lr=1e-4, capturable=True)
        do = S.Adam(da.parameters(), lr=5e-5, capturable=True)
        pol = S.MultiGeneratorPolicy(S.PolicyConfig(num_generators=1, force=S.GAN))
        return ga, da, S.MultiGeneratorGAN([ga], [go], crit, discriminator=da, d_optimizer=do, policy=pol,
                                           loss_allreduce=S.parallel.mean_over_ranks(), use_cuda_graphs=graphs)
    ga, da, ta = make(False)
    gb, db, tb = make(True)
    same = True
    for _ in range(3):
        la = ta.step(lr, hr).clone()
        lb = tb.step(lr, hr).clone()
        torch.cuda.synchronize()
        same = same and torch.equal(la, lb)
    same = same and torch.equal(ga.flat_parameters(), gb.flat_parameters()) and torch.equal(da.flat_parameters(), db.flat_parameters())
    flag = torch.tensor([1 if same else 0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print(f"adversarial steps from CUDA graphs under data parallelism == eager on every rank: {bool(flag.item())}")
        ok = ok and bool(flag.item())
    return ok


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    if "adversarial" in sys.argv[1:]:
        ok = check_adversarial(rank, world, dev)
        if rank == 0:
            print("MULTIGPU ADVERSARIAL CHECK", "PASS" if ok else "FAIL")
        S.parallel.shutdown_nccl()
        dist.barrier()
        torch.cuda.synchronize()
        sys.stdout.flush()
        os._exit(0 if ok else 1)
    fused = 1 if "fused" in sys.argv[1:] else (0 if "perlayer" in sys.argv[1:] else 2)
    S.lib().srg_set_trunk_fused(fused)
    b, H, W = 4, 24, 40
    torch.manual_seed(0)
    lr_full = torch.rand(world * b, 3, H, W)
    hr_full = torch.rand(world * b, 3, 4 * H, 4 * W)
    torch.manual_seed(7 + rank)                       # different init per rank: data_parallel must broadcast rank 0's
    g = S.SRResNet(num_residuals=3).to(dev)
    transport = "nccl" if "nccl" in sys.argv[1:] else "peer"
    S.parallel.data_parallel([g], sync_batchnorm=True, sync_bn_transport=transport)
    crit = S.ReconstructionLoss()
    lr = S.parallel.shard_batch(lr_full, rank, world).to(dev)
    hr = S.parallel.shard_batch(hr_full, rank, world).to(dev)
    sr = g(lr)
    com, tv = crit(hr, sr)
    (com + tv).backward()
    torch.cuda.synchronize()
    flat_dp = g.flat_grads().clone()
    rm_dp = g.state_dict()["residual_blocks.0.bn1.running_mean"].clone()
    rv_dp = g.state_dict()["residual_blocks.2.bn2.running_var"].clone()
    # all ranks must hold identical gradients after the all-reduce
    ref = flat_dp.clone()
    dist.broadcast(ref, src=0)
    same = float((ref - flat_dp).abs().max())
    ok = True
    if rank == 0:
        torch.manual_seed(7)
        g1 = S.SRResNet(num_residuals=3).to(dev)      # same init as rank 0's model, no SyncBN, whole batch
        sr1 = g1(lr_full.to(dev))
        loss = 0
        for r in range(world):
            c, t = crit(hr_full[r * b:(r + 1) * b].to(dev), sr1[r * b:(r + 1) * b])
            loss = loss + (c + t) / world
        loss.backward()
        torch.cuda.synchronize()
        f1 = g1.flat_grads()
        err = float((flat_dp - f1).abs().max() / f1.abs().max())
        l2 = float((flat_dp - f1).norm() / f1.norm())
        worst, worst_name = 0.0, ""
        for (name, off, n, shape) in g._ptable:
            a_, b_ = flat_dp[off:off + n], f1[off:off + n]
            if float(b_.abs().max()) < 1e-12:
                continue
            r_ = float((a_ - b_).norm() / b_.norm())
            if r_ > worst:
                worst, worst_name = r_, name
        print(f"worst per-tensor grad l2-rel {worst:.3e} ({worst_name})")
        e_rm = float((rm_dp - g1.state_dict()["residual_blocks.0.bn1.running_mean"]).abs().max())
        e_rv = float((rv_dp - g1.state_dict()["residual_blocks.2.bn2.running_var"]).abs().max())
        e_sr = float((sr.detach() - sr1.detach()[:b]).abs().max() / sr1.detach().abs().max())
        eng = g.last_engine()
        print(f"transport={transport} peer_sync_errors={S.parallel.peer_sync_errors()} trunk layers per launch "
              f"{S.lib().srg_generator_trunk_layers(eng.handle)} trunk error word {S.lib().srg_generator_trunk_error(eng.handle)}")
        print(f"world={world}: DP vs single-process global batch: SR max-rel {e_sr:.3e}; grads max-rel {err:.3e} l2-rel {l2:.3e}; "
              f"running_mean diff {e_rm:.3e}; running_var diff {e_rv:.3e}; cross-rank grad diff {same:.3e}")
        ok = e_sr < 1e-2 and l2 < 5e-2 and worst < 5e-2 and e_rm < 1e-4 and e_rv < 1e-4 and same == 0.0
        print("MULTIGPU CHECK", "PASS" if ok else "FAIL")
    S.parallel.shutdown_nccl()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
