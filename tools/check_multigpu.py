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


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    b, H, W = 4, 24, 40
    torch.manual_seed(0)
    lr_full = torch.rand(world * b, 3, H, W)
    hr_full = torch.rand(world * b, 3, 4 * H, 4 * W)
    torch.manual_seed(7 + rank)                       # different init per rank: data_parallel must broadcast rank 0's
    g = S.SRResNet(num_residuals=3).to(dev)
    transport = sys.argv[1] if len(sys.argv) > 1 else "peer"
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
        print(f"transport={transport} peer_sync_errors={S.parallel.peer_sync_errors()}")
        print(f"world={world}: DP vs single-process global batch: SR max-rel {e_sr:.3e}; grads max-rel {err:.3e} l2-rel {l2:.3e}; "
              f"running_mean diff {e_rm:.3e}; running_var diff {e_rv:.3e}; cross-rank grad diff {same:.3e}")
        ok = e_sr < 1e-2 and l2 < 5e-2 and worst < 5e-2 and e_rm < 1e-4 and e_rv < 1e-4 and same == 0.0
        print("MULTIGPU CHECK", "PASS" if ok else "FAIL")
    S.parallel.shutdown_nccl()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
