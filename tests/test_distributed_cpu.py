"""world_size-2 gloo tests (CPU) of the host-side data-parallel logic: batch sharding, gradient averaging hook,
loss averaging for the policy, and identical policy decisions on all ranks."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import srgan_b200 as S
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # 1. sharding: contiguous, disjoint, covering
        full = torch.arange(8 * 3).view(8, 3)
        mine = S.parallel.shard_batch(full, rank, world)
        gathered = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(gathered, mine)
        assert torch.equal(torch.cat(gathered), full)
        # 2. gradient hook = mean over ranks of the flat buffer (DDP semantics)
        flat = torch.full((1000,), float(rank + 1))
        S.parallel.average_gradients_hook()(None, flat)
        assert torch.allclose(flat, torch.full((1000,), (1 + world) / 2.0))
        # 3. policy fed with rank-averaged losses takes identical decisions everywhere
        pol = S.MultiGeneratorPolicy(S.PolicyConfig(num_generators=3, seed=3, starting_gan_loss=0.2))
        mean = S.parallel.mean_over_ranks()
        decisions = []
        for t in range(6):
            plan = pol.plan_batch()
            decisions.append(plan)
            local = torch.tensor([[0.0, 0.3 / (t + 1) + 0.01 * rank + 0.05 * gid, 0.0, 0.0] for gid, _ in plan])
            mean(local)
            for (gid, _), row in zip(plan, local.tolist()):
                pol.observe(gid, row[1])
        order = pol.end_epoch()
        blob = [None] * world
        dist.all_gather_object(blob, (decisions, order, pol.running))
        assert all(b == blob[0] for b in blob)
        q.put((rank, "ok"))
    except Exception as e:  # pragma: no cover
        q.put((rank, repr(e)))
    finally:
        dist.destroy_process_group()


def test_world2_gloo_host_logic():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, "ok"), (1, "ok")], res


def test_shard_batch_rejects_ragged():
    import srgan_b200 as S
    with pytest.raises(ValueError):
        S.parallel.shard_batch(torch.zeros(10, 3), 0, 4)
