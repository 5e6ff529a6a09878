"""Single-node data parallelism: one process per GPU, batch sharded across ranks (reference: DDP over NCCL,
src/train.py:29-35,45-47).

Collectives on the path:
  * per-model gradient all-reduce (mean) over ONE flat fp32 buffer (6.2 MB per generator, 11.1 MB discriminator),
    enqueued right after the model's backward kernels via ``set_grad_hook``;
  * SyncBatchNorm statistics: per-channel (sum, sum of squares) forward and (sum dy, sum dy*y) backward, 128 doubles
    per BatchNorm layer, all-reduced between the local reduction kernel and the finalize kernel on the same stream.
``torch.distributed`` supplies rendezvous and the process group; every model gets its own NCCL communicator created
from the library (``srg_nccl_comm_create``), so the engine can enqueue the SyncBN collectives from C++ between its own
kernels, the collectives are captured into the step's CUDA graph, and the K generators' branches never share one.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import c_void_p
from typing import Iterable, Optional

import torch
import torch.distributed as dist

from . import _lib
from ._lib import check, stream_ptr

_comms = []          # communicator handles created by this process (destroyed by shutdown_nccl)


def shard_batch(t: torch.Tensor, rank: int, world: int) -> torch.Tensor:
    """Rank's contiguous slice of a global batch (the reference's DistributedSampler gives each rank N/world samples)."""
    n = t.shape[0]
    if n % world != 0:
        raise ValueError(f"global batch {n} is not divisible by world size {world}")
    per = n // world
    return t[rank * per:(rank + 1) * per]


def create_comm(group: Optional[dist.ProcessGroup] = None) -> c_void_p:
    """A new library-side NCCL communicator over the ranks of ``group`` (rendezvous through torch.distributed).
    Every model gets its own so that the models' steps can run concurrently (separate streams / graph branches)."""
    L = _lib.lib()
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    uid = torch.zeros(128, dtype=torch.uint8)
    if rank == 0:
        buf = ctypes.create_string_buffer(128)
        check(L.srg_nccl_unique_id(buf), "srg_nccl_unique_id")
        uid = torch.frombuffer(bytearray(buf.raw), dtype=torch.uint8).clone()
    if dist.get_backend(group) == "nccl":
        dev_uid = uid.cuda()
        dist.broadcast(dev_uid, src=0, group=group)
        uid = dev_uid.cpu()
    else:
        dist.broadcast(uid, src=0, group=group)
    comm = c_void_p()
    check(L.srg_nccl_comm_create(bytes(uid.tolist()), world, rank, ctypes.byref(comm)), "srg_nccl_comm_create")
    _comms.append(comm)
    return comm


_peers = []


def create_peer_sync(group: Optional[dist.ProcessGroup] = None) -> c_void_p:
    """NVLink peer-memory exchange object for SyncBatchNorm (csrc/peer_sync.cu): every rank allocates an exchange
    buffer, the CUDA IPC handles are all-gathered through torch.distributed and each rank maps its peers' buffers."""
    L = _lib.lib()
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    ps = c_void_p()
    check(L.srg_peer_sync_create(ctypes.byref(ps), world, rank), "srg_peer_sync_create")
    buf = ctypes.create_string_buffer(64)
    check(L.srg_peer_sync_handle(ps, buf), "srg_peer_sync_handle")
    mine = torch.frombuffer(bytearray(buf.raw), dtype=torch.uint8).clone()
    on_gpu = dist.get_backend(group) == "nccl"
    if on_gpu:
        mine = mine.cuda()
    gathered = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(gathered, mine, group=group)
    blob = b"".join(bytes(t.cpu().tolist()) for t in gathered)
    check(L.srg_peer_sync_connect(ps, blob), "srg_peer_sync_connect")
    dist.barrier(group=group)
    _peers.append(ps)
    return ps


def peer_sync_errors() -> int:
    """Number of peer-sync objects on which a wait for a peer ever timed out (0 in a healthy run)."""
    L = _lib.lib()
    return sum(int(L.srg_peer_sync_error(p)) for p in _peers)


def check_peer_sync() -> None:
    """Raise if a SyncBatchNorm peer exchange ever gave up waiting for a peer GPU (csrc/peer_sync.cu poisons the
    statistics of that step with NaN; this turns it into an exception at the next loss read-back).  No-op without
    peer-sync objects; otherwise one 4-byte device read per object, so call it where the host synchronises anyway."""
    if _peers and peer_sync_errors():
        raise RuntimeError("SyncBatchNorm peer exchange timed out: a peer GPU did not deliver its BatchNorm statistics "
                           "(rank stalled or lost); the affected steps carry NaN statistics")


def shutdown_nccl() -> None:
    L = _lib.lib()
    while _comms:
        L.srg_nccl_comm_destroy(_comms.pop())


def average_gradients_hook(group: Optional[dist.ProcessGroup] = None, comm: Optional[c_void_p] = None):
    """Returns hook(module, flat_grads) that all-reduces the flat gradient buffer (mean over ranks), stream-ordered
    after the backward kernels -- the DDP gradient all-reduce of src/train.py:195 in ONE collective.  With ``comm``
    (create_comm) the collective goes through the model's own communicator on the current stream (graph-capturable,
    concurrent with other models); without it through torch.distributed (gloo on CPU in the tests)."""
    def hook(module, flat_grads: torch.Tensor) -> None:
        world = dist.get_world_size(group)
        if world == 1:
            return
        if os.environ.get("SRG_DP_NO_ALLREDUCE") == "1":
            return          # MEASUREMENT ONLY (bench.py apportions the multi-GPU step): ranks drift apart, never train like this
        if comm is not None:
            # ncclAvg: sum and 1/world in the collective itself, no extra pass over the gradients
            check(_lib.lib().srg_nccl_allreduce_mean_f32(comm, c_void_p(flat_grads.data_ptr()), flat_grads.numel(),
                                                         stream_ptr()), "srg_nccl_allreduce_mean_f32")
        else:
            dist.all_reduce(flat_grads, op=dist.ReduceOp.SUM, group=group)     # gloo (CPU tests) has no averaging op
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


def data_parallel(modules: Iterable, group: Optional[dist.ProcessGroup] = None, sync_batchnorm: bool = True,
                  sync_bn_transport: str = "peer") -> None:
    """Make ``modules`` (SRResNet / Discriminator instances, already on their device) train data-parallel over
    ``group``: broadcast rank 0's state, give every module its own NCCL communicator, install the flat-gradient
    all-reduce hook and (generators) SyncBatchNorm on that communicator."""
    modules = list(modules)
    broadcast_parameters(modules, group)
    world = dist.get_world_size(group)
    for m in modules:
        comm = create_comm(group) if (world > 1 and dist.get_backend(group) == "nccl") else None
        m.set_grad_hook(average_gradients_hook(group, comm))
        if sync_batchnorm and comm is not None and hasattr(m, "enable_sync_batchnorm"):
            if sync_bn_transport == "peer":
                m.enable_sync_batchnorm(world=world, peer_sync=create_peer_sync(group))
            else:
                m.enable_sync_batchnorm(comm=comm, world=world)


def mean_over_ranks(group: Optional[dist.ProcessGroup] = None):
    """callable(tensor): in-place mean over ranks (keeps the multi-generator policy identical on every rank)."""
    def fn(t: torch.Tensor) -> None:
        world = dist.get_world_size(group)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
            t.mul_(1.0 / world)
    return fn
