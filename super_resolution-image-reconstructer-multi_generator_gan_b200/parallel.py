"""Single-node data parallelism: one process per GPU, batch sharded across ranks (reference: DDP over NCCL,
src/train.py:29-35,45-47).

Collectives on the path:
  * per-model gradient all-reduce (mean) over ONE flat fp32 buffer (6.2 MB per generator, 11.1 MB discriminator),
    enqueued right after the model's backward kernels via ``set_grad_hook``;
  * SyncBatchNorm statistics: per-channel (sum, sum of squares) forward and (sum dy, sum dy*y) backward, 128 doubles
    per BatchNorm layer, all-reduced between the local reduction kernel and the finalize kernel on the same stream.
``torch.distributed`` supplies rendezvous and the process group; the SyncBN hook uses a dedicated NCCL communicator
created from the library (``srg_nccl_init``) so the engine can enqueue it from C++ between its own kernels.
"""
from __future__ import annotations

import ctypes
from ctypes import c_void_p
from typing import Iterable, Optional

import torch
import torch.distributed as dist

from . import _lib
from ._lib import check, stream_ptr

_nccl_ready = False


def shard_batch(t: torch.Tensor, rank: int, world: int) -> torch.Tensor:
    """Rank's contiguous slice of a global batch (the reference's DistributedSampler gives each rank N/world samples)."""
    n = t.shape[0]
    if n % world != 0:
        raise ValueError(f"global batch {n} is not divisible by world size {world}")
    per = n // world
    return t[rank * per:(rank + 1) * per]


def init_nccl(group: Optional[dist.ProcessGroup] = None) -> None:
    """Create the library-side NCCL communicator (used by SyncBatchNorm) over the ranks of ``group``."""
    global _nccl_ready
    if _nccl_ready:
        return
    L = _lib.lib()
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    uid = torch.zeros(128, dtype=torch.uint8)
    if rank == 0:
        buf = ctypes.create_string_buffer(128)
        check(L.srg_nccl_unique_id(buf), "srg_nccl_unique_id")
        uid = torch.frombuffer(bytearray(buf.raw), dtype=torch.uint8).clone()
    backend = dist.get_backend(group)
    if backend == "nccl":
        dev_uid = uid.cuda()
        dist.broadcast(dev_uid, src=0, group=group)
        uid = dev_uid.cpu()
    else:
        dist.broadcast(uid, src=0, group=group)
    raw = bytes(uid.tolist())
    check(L.srg_nccl_init(raw, world, rank), "srg_nccl_init")
    _nccl_ready = True


def shutdown_nccl() -> None:
    global _nccl_ready
    if _nccl_ready:
        _lib.lib().srg_nccl_shutdown()
        _nccl_ready = False


def average_gradients_hook(group: Optional[dist.ProcessGroup] = None):
    """Returns hook(module, flat_grads) that all-reduces the flat gradient buffer (mean over ranks), stream-ordered
    after the backward kernels -- the DDP gradient all-reduce of src/train.py:195 in one collective."""
    def hook(module, flat_grads: torch.Tensor) -> None:
        world = dist.get_world_size(group)
        if world == 1:
            return
        dist.all_reduce(flat_grads, op=dist.ReduceOp.SUM, group=group)
        flat_grads.mul_(1.0 / world)
    return hook


def broadcast_parameters(modules: Iterable, group: Optional[dist.ProcessGroup] = None, src: int = 0) -> None:
    """DDP-constructor semantics (src/train.py:45,47): rank 0's parameters and buffers to every rank."""
    for m in modules:
        flat = m.flat_parameters()
        dist.broadcast(flat, src=src, group=group)
        fb = m._rt.get("flat_buf")
        if fb is not None:
            dist.broadcast(fb, src=src, group=group)


def data_parallel(modules: Iterable, group: Optional[dist.ProcessGroup] = None, sync_batchnorm: bool = True) -> None:
    """Make ``modules`` (SRResNet / Discriminator instances, already on their device) train data-parallel over
    ``group``: broadcast rank 0's state, install the gradient all-reduce hook and (generators) SyncBatchNorm."""
    modules = list(modules)
    broadcast_parameters(modules, group)
    hook = average_gradients_hook(group)
    for m in modules:
        m.set_grad_hook(hook)
        if sync_batchnorm and hasattr(m, "enable_sync_batchnorm") and dist.get_world_size(group) > 1:
            init_nccl(group)
            m.enable_sync_batchnorm()


def mean_over_ranks(group: Optional[dist.ProcessGroup] = None):
    """callable(tensor): in-place mean over ranks (keeps the multi-generator policy identical on every rank)."""
    def fn(t: torch.Tensor) -> None:
        world = dist.get_world_size(group)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
            t.mul_(1.0 / world)
    return fn
