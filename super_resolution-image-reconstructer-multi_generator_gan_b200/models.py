"""nn.Module surface of the reference's models (reference: src/models.py) over libsrgan_b200.so.

``SRResNet`` keeps the reference's constructor signature, sub-module tree, registration order, default
initialisation (so ``torch.manual_seed(s); SRResNet()`` yields the reference's weights) and ``state_dict`` keys
(SURVEY Appendix A), but none of its arithmetic: ``forward`` / ``backward`` are single calls into the C ABI
(``srg_generator_forward`` / ``srg_generator_backward``), which enqueue hand-written sm_100a kernels on the current
CUDA stream.  There is no PyTorch / cuDNN / CPU fallback: the sub-modules below only *hold* parameters.

Parameter storage: every ``nn.Parameter`` is a view into ONE flat fp32 buffer laid out as the engine's parameter table
(``srg_generator_param_info``); gradients come back as views of one flat buffer of the same layout, so the optimiser
(``optim.Adam``) and the data-parallel gradient all-reduce are single launches over contiguous memory.
"""
from __future__ import annotations

import ctypes
import itertools
import math
import weakref
from ctypes import byref, c_char_p, c_int, c_int64, c_void_p, create_string_buffer
from typing import Dict, List, Optional, Tuple

import torch
import torch.nn as nn

from . import _lib
from ._flat import _FlatModule
from ._lib import check, stream_ptr


# ----------------------------------------------------------------------------------------------------------------
# parameter holders (same registration order / default init as nn.Conv2d / nn.BatchNorm2d; never executed)
# ----------------------------------------------------------------------------------------------------------------
class _Holder(nn.Module):
    def forward(self, *a, **k):  # pragma: no cover - guard
        raise RuntimeError(
            f"{type(self).__name__} only holds parameters; the arithmetic lives in libsrgan_b200.so and is reached "
            "through the owning SRResNet / Discriminator module (no per-layer PyTorch fallback exists)")


class ConvParams(_Holder):
    """weight [cout, cin, k, k] + bias [cout] with nn.Conv2d's default initialisation
    (kaiming_uniform(a=sqrt(5)) then U(+-1/sqrt(fan_in)) for the bias, drawn in that order)."""

    def __init__(self, cin: int, cout: int, kernel_size: int, stride: int = 1, padding: int = 0):
        super().__init__()
        self.in_channels, self.out_channels = cin, cout
        self.kernel_size, self.stride, self.padding = kernel_size, stride, padding
        self.weight = nn.Parameter(torch.empty(cout, cin, kernel_size, kernel_size))
        self.bias = nn.Parameter(torch.empty(cout))
        nn.init.kaiming_uniform_(self.weight, a=math.sqrt(5))
        bound = 1.0 / math.sqrt(cin * kernel_size * kernel_size)
        nn.init.uniform_(self.bias, -bound, bound)

    def extra_repr(self):
        return f"{self.in_channels}, {self.out_channels}, kernel_size={self.kernel_size}, stride={self.stride}, padding={self.padding}"


class BatchNormParams(_Holder):
    """nn.BatchNorm2d(C) state: weight=1, bias=0, running_mean=0, running_var=1, num_batches_tracked=0."""

    def __init__(self, channels: int, eps: float = 1e-5, momentum: float = 0.1):
        super().__init__()
        self.num_features, self.eps, self.momentum = channels, eps, momentum
        self.weight = nn.Parameter(torch.ones(channels))
        self.bias = nn.Parameter(torch.zeros(channels))
        self.register_buffer("running_mean", torch.zeros(channels))
        self.register_buffer("running_var", torch.ones(channels))
        self.register_buffer("num_batches_tracked", torch.tensor(0, dtype=torch.long))


class _Stateless(_Holder):
    """Placeholder keeping nn.Sequential indices aligned with the reference (PixelShuffle / ReLU / ... slots)."""

    def __init__(self, what: str):
        super().__init__()
        self.what = what

    def extra_repr(self):
        return self.what


class ResidualBlock(_Holder):
    """Parameter tree of the reference's ResidualBlock (src/models.py:10-25): conv1, bn1, relu, conv2, bn2.
    Executed only as part of SRResNet (the engine fuses the whole trunk)."""

    def __init__(self, num_features: int):
        super().__init__()
        self.conv1 = ConvParams(num_features, num_features, 3, padding=1)
        self.bn1 = BatchNormParams(num_features)
        self.relu = _Stateless("ReLU(inplace=True)")
        self.conv2 = ConvParams(num_features, num_features, 3, padding=1)
        self.bn2 = BatchNormParams(num_features)


# ---- engine ownership -----------------------------------------------------------------------------------------
# A grad-enabled forward marks its engine busy until backward has run.  A forward whose output is dropped WITHOUT a
# backward would otherwise keep the engine (and its multi-GB training workspace) for ever and make every such call
# allocate a new one: the autograd node's death releases it.  The token keeps a late finalizer of an OLD node (nodes
# also die after their backward) from releasing an engine that a newer forward has already re-acquired.
_busy_tokens = itertools.count(2)          # never 1: `True == 1`, and train.py marks engines with plain True


def _acquire(eng) -> int:
    eng.busy = next(_busy_tokens)
    return eng.busy


def _release_if(eng, token: int) -> None:
    if eng.busy == token:
        eng.busy = False


def _release_on_death(ctx, eng) -> None:
    ctx.busy_token = eng.busy
    weakref.finalize(ctx, _release_if, eng, eng.busy)


# ----------------------------------------------------------------------------------------------------------------
# engine wrapper
# ----------------------------------------------------------------------------------------------------------------
class _GeneratorEngine:
    """One srg_generator_t bound to a workspace: fixed (N, H, W) geometry, one in-flight forward at a time."""

    def __init__(self, N: int, H: int, W: int, n_res: int, n_up: int, training: bool, device: torch.device):
        L = _lib.lib()
        self.handle = c_void_p()
        check(L.srg_generator_create(byref(self.handle), N, H, W, n_res, n_up), "srg_generator_create")
        self.N, self.H, self.W, self.n_res, self.n_up, self.training = N, H, W, n_res, n_up, training
        self.device = device
        self.busy = False          # a training forward whose backward has not run yet
        self.ws = None
        self.bound_key = None
        self.param_elems = int(L.srg_generator_param_elems(self.handle))
        self.buffer_elems = int(L.srg_generator_buffer_elems(self.handle))

    def __del__(self):
        try:
            if self.handle:
                _lib.lib().srg_generator_destroy(self.handle)
                self.handle = None
        except Exception:
            pass

    def param_table(self) -> List[Tuple[str, int, int, Tuple[int, ...]]]:
        L = _lib.lib()
        out = []
        name = create_string_buffer(128)
        off, numel, ndim = c_int64(), c_int64(), c_int()
        shape = (c_int * 4)()
        for i in range(L.srg_generator_num_params(self.handle)):
            check(L.srg_generator_param_info(self.handle, i, name, 128, byref(off), byref(numel), byref(ndim), shape))
            out.append((name.value.decode(), off.value, numel.value, tuple(shape[k] for k in range(ndim.value))))
        return out

    def buffer_table(self) -> List[Tuple[str, int, int]]:
        L = _lib.lib()
        out = []
        name = create_string_buffer(128)
        off, numel = c_int64(), c_int64()
        for i in range(L.srg_generator_num_buffers(self.handle)):
            check(L.srg_generator_buffer_info(self.handle, i, name, 128, byref(off), byref(numel)))
            out.append((name.value.decode(), off.value, numel.value))
        return out

    def tensor_table(self) -> Dict[str, Tuple[int, Tuple[int, ...], int]]:
        L = _lib.lib()
        out = {}
        name = create_string_buffer(128)
        off, dt = c_int64(), c_int()
        dims = (c_int * 4)()
        for i in range(L.srg_generator_num_tensors(self.handle)):
            check(L.srg_generator_tensor_info(self.handle, i, name, 128, byref(off), dims, byref(dt)))
            out[name.value.decode()] = (off.value, tuple(dims[k] for k in range(4)), dt.value)
        return out

    def bind(self, flat_params: torch.Tensor, flat_grads: Optional[torch.Tensor], flat_buffers: torch.Tensor):
        L = _lib.lib()
        key = (flat_params.data_ptr(), flat_buffers.data_ptr())
        if self.ws is None:
            nbytes = int(L.srg_generator_workspace_bytes(self.handle, 1 if self.training else 0))
            self.ws = torch.empty(nbytes + 1024, dtype=torch.uint8, device=self.device)
        if key != self.bound_key:
            base = self.ws.data_ptr()
            aligned = (base + 1023) & ~1023
            check(L.srg_generator_bind(self.handle, c_void_p(flat_params.data_ptr()),
                                       c_void_p(flat_grads.data_ptr()) if flat_grads is not None else None,
                                       c_void_p(flat_buffers.data_ptr()), c_void_p(aligned),
                                       self.ws.numel() - (aligned - base), 1 if self.training else 0),
                  "srg_generator_bind")
            self.ws_base = aligned
            self.bound_key = key
        elif flat_grads is not None:
            check(L.srg_generator_set_grads(self.handle, c_void_p(flat_grads.data_ptr())))

    def named_tensor(self, name: str) -> torch.Tensor:
        """bf16 NHWC intermediate inside the workspace (per-layer parity checks)."""
        off, dims, dt = self.tensor_table()[name]
        n = dims[0] * dims[1] * dims[2] * dims[3]
        start = (self.ws_base - self.ws.data_ptr()) + off
        return self.ws[start:start + 2 * n].view(torch.bfloat16).view(*dims)


class _GeneratorFn(torch.autograd.Function):
    """autograd node for one SRResNet pass: forward = srg_generator_forward, backward = srg_generator_backward."""

    @staticmethod
    def forward(ctx, module, eng, lr_imgs, *params):
        L = _lib.lib()
        sr = torch.empty(eng.N, 3, eng.H << eng.n_up, eng.W << eng.n_up, dtype=torch.float32, device=lr_imgs.device)
        check(L.srg_generator_forward(eng.handle, c_void_p(lr_imgs.data_ptr()), c_void_p(sr.data_ptr()),
                                      1 if module.training else 0, 1 if module.training else 0, stream_ptr()),
              "srg_generator_forward")
        ctx.module_ref = weakref.ref(module)
        ctx.eng = eng
        _release_on_death(ctx, eng)
        ctx.n_params = len(params)
        ctx.set_materialize_grads(False)
        return sr

    @staticmethod
    def backward(ctx, dsr):
        module = ctx.module_ref()
        eng = ctx.eng
        if dsr is None or module is None:
            _release_if(eng, ctx.busy_token)
            return (None,) * (3 + ctx.n_params)
        L = _lib.lib()
        dsr = dsr.contiguous()
        if dsr.dtype != torch.float32:
            dsr = dsr.float()
        flat_g = module._grad_buffer_for_backward(eng)
        check(L.srg_generator_set_grads(eng.handle, c_void_p(flat_g.data_ptr())))
        check(L.srg_generator_backward(eng.handle, c_void_p(dsr.data_ptr()), stream_ptr()), "srg_generator_backward")
        _release_if(eng, ctx.busy_token)
        module._after_backward(flat_g)
        grads = tuple(flat_g[off:off + n].view(shape) for (_, off, n, shape) in module._ptable)
        return (None, None, None) + grads


class SRResNet(_FlatModule):
    """Drop-in for the reference's SRResNet (src/models.py:44-87).

    conv 9x9 (3->64) + LeakyReLU(0.2) -> ``num_residuals`` x [conv3x3, BN, ReLU, conv3x3, BN, +skip] -> conv3x3 ->
    + global skip -> ``int(upscale_factor // 2)`` x [conv3x3 (64->256), PixelShuffle(2), ReLU] -> conv 9x9 (64->3).
    Input / output: NCHW fp32 CUDA tensors.  The kernels are specialised for in_channels=3, num_features=64 (the
    reference's only configuration); other widths raise.
    """

    def __init__(self, in_channels: int = 3, num_features: int = 64, num_residuals: int = 16, upscale_factor: int = 4):
        super().__init__()
        if in_channels != 3 or num_features != 64:
            raise NotImplementedError("libsrgan_b200 implements the reference configuration in_channels=3, num_features=64")
        self.in_channels, self.num_features = in_channels, num_features
        self.num_residuals = num_residuals
        self.num_upsample_stages = int(upscale_factor // 2)      # reference quirk: 2->1, 3->1, 4->2, 8->4 stages
        self.conv1 = ConvParams(in_channels, num_features, 9, padding=4)
        self.relu = _Stateless("LeakyReLU(0.2, inplace=True)")
        self.residual_blocks = nn.Sequential(*[ResidualBlock(num_features) for _ in range(num_residuals)])
        self.conv2 = ConvParams(num_features, num_features, 3, padding=1)
        ups: List[nn.Module] = []
        for _ in range(self.num_upsample_stages):
            ups += [ConvParams(num_features, num_features * 4, 3, padding=1), _Stateless("PixelShuffle(2)"),
                    _Stateless("ReLU(inplace=True)")]
        self.upsample = nn.Sequential(*ups)
        self.conv3 = ConvParams(num_features, in_channels, 9, padding=4)
        self._reset_runtime()

    def _probe_engine(self, device) -> _GeneratorEngine:
        rt = self._rt
        if rt.get("probe") is None:
            rt["probe"] = _GeneratorEngine(1, 8, 8, self.num_residuals, self.num_upsample_stages, False, device)
        return rt["probe"]

    def _tables(self, device):
        probe = self._probe_engine(device)
        return probe.param_table(), probe.buffer_table(), probe.param_elems, probe.buffer_elems

    # ---- engines ---------------------------------------------------------------------------------------------------
    def _engine(self, N: int, H: int, W: int, training: bool, device, need_grad: bool) -> _GeneratorEngine:
        rt = self._rt
        key = (N, H, W, training)
        pool = rt["engines"].setdefault(key, [])
        eng = None
        for e in pool:
            if not e.busy:
                eng = e
                break
        if eng is None:
            eng = _GeneratorEngine(N, H, W, self.num_residuals, self.num_upsample_stages, training, device)
            if rt["sync_bn"]:
                self._attach_sync_bn(eng)
            if getattr(self, "debug_keep_grads", False):
                check(_lib.lib().srg_generator_set_keep_grads(eng.handle, 1))
            if rt.get("profile"):
                check(_lib.lib().srg_generator_profile_enable(eng.handle, 1))
            pool.append(eng)
        return eng

    def _attach_sync_bn(self, eng):
        kind, obj, world = self._rt["sync_bn"]
        if kind == "peer":
            check(_lib.lib().srg_generator_use_peer_sync(eng.handle, obj), "srg_generator_use_peer_sync")
        else:
            check(_lib.lib().srg_generator_use_nccl(eng.handle, obj, world), "srg_generator_use_nccl")

    def enable_sync_batchnorm(self, comm=None, world: int = 1, peer_sync=None):
        """SyncBatchNorm: per-channel sums are reduced across ranks between the local reduction and the BatchNorm
        finalize, forward and backward -- through ``peer_sync`` (parallel.create_peer_sync(): NVLink peer-memory
        exchange fused with the finalize kernel) or an NCCL communicator (parallel.create_comm())."""
        self._rt["sync_bn"] = ("peer", peer_sync, int(world)) if peer_sync is not None else ("nccl", comm, int(world))
        for pool in self._rt["engines"].values():
            for e in pool:
                self._attach_sync_bn(e)

    def launch_count(self) -> int:
        L = _lib.lib()
        return sum(int(L.srg_generator_launch_count(e.handle)) for pool in self._rt["engines"].values() for e in pool)

    def profile_enable(self, on: bool = True) -> None:
        """CUDA-event timing of the dominant kernel class (3x3 64->64 conv fprop/dgrad launches), see bench.py."""
        self._rt["profile"] = bool(on)
        for pool in self._rt["engines"].values():
            for e in pool:
                check(_lib.lib().srg_generator_profile_enable(e.handle, 1 if on else 0))

    def profile_read(self) -> Tuple[float, int]:
        """(summed device ms, launches) of the profiled kernel class since the last read."""
        from ctypes import c_double, c_longlong
        total, count = 0.0, 0
        for pool in self._rt["engines"].values():
            for e in pool:
                ms, n = c_double(), c_longlong()
                check(_lib.lib().srg_generator_profile_read(e.handle, byref(ms), byref(n)), "profile_read")
                total += ms.value
                count += n.value
        return total, count

    def last_engine(self) -> Optional[_GeneratorEngine]:
        return self._rt["last_engine"]

    #: eval-mode activation budget in bytes: larger batches are streamed in chunks of frames (None: never chunk)
    stream_budget_bytes: Optional[int] = 6 << 30

    def _stream_chunk(self, N: int, H: int, W: int, device) -> int:
        """Frames per eval-mode engine call such that the engine workspace stays within ``stream_budget_bytes``."""
        budget = self.stream_budget_bytes
        if budget is None:
            return N
        rt = self._rt
        key = ("ws1", H, W)
        if key not in rt:
            probe = _GeneratorEngine(1, H, W, self.num_residuals, self.num_upsample_stages, False, device)
            rt[key] = int(_lib.lib().srg_generator_workspace_bytes(probe.handle, 0))
            del probe
        per_frame = max(rt[key], 1)
        return max(1, min(N, int(budget // per_frame)))

    # ---- forward ---------------------------------------------------------------------------------------------------
    def forward(self, x: torch.Tensor) -> torch.Tensor:
        if not x.is_cuda:
            raise RuntimeError("SRResNet (libsrgan_b200) runs on a CUDA device only; there is no CPU path")
        if x.dim() != 4 or x.shape[1] != 3:
            raise RuntimeError(f"expected an N x 3 x H x W input, got {tuple(x.shape)}")
        x = x.contiguous()
        if x.dtype != torch.float32:
            x = x.float()
        N, _, H, W = x.shape
        self._flatten(x.device)
        rt = self._rt
        need_grad = self.training and torch.is_grad_enabled()
        if not self.training and N > 1:
            # Evaluation (src/evaluation.py:48-50, BASELINE configs[3]): frames are independent in eval mode (running
            # BatchNorm statistics), so a batch whose activations would not fit the streaming budget is run in chunks
            # of whole frames through ONE smaller engine: peak memory is O(chunk), the arithmetic per frame (and hence
            # every output bit) is that of the unchunked call.  8 x 1080p would otherwise materialise ~34 GB.
            chunk = self._stream_chunk(N, H, W, x.device)
            if chunk < N:
                sr = torch.empty(N, 3, H << self.num_upsample_stages, W << self.num_upsample_stages, dtype=torch.float32,
                                 device=x.device)
                packed = set()                             # engines whose bf16 weight copy is current for THIS call
                for i in range(0, N, chunk):
                    j = min(N, i + chunk)
                    self._eval_into(x[i:j], sr[i:j], packed)   # whole frames of an NCHW batch are contiguous: no copy
                return sr
        eng = self._engine(N, H, W, self.training, x.device, need_grad)
        if self.training and getattr(eng, "grad_flat", None) is None:
            eng.grad_flat = torch.empty_like(rt["flat"])
        eng.bind(rt["flat"], eng.grad_flat if self.training else None, rt["flat_buf"])
        L = _lib.lib()
        check(L.srg_generator_pack(eng.handle, stream_ptr()), "srg_generator_pack")
        rt["last_engine"] = eng
        if self.training and rt["nbt"] is not None and self.num_residuals > 0:
            rt["nbt"] += 1
        if need_grad:
            _acquire(eng)
            return _GeneratorFn.apply(self, eng, x, *rt["plist"])
        # eval mode (running statistics) or no_grad: no autograd graph.  The reference keeps a graph through an
        # eval-mode generator in train_discriminator (src/train.py:212) but discards those gradients.
        sr = torch.empty(N, 3, H << self.num_upsample_stages, W << self.num_upsample_stages, dtype=torch.float32,
                         device=x.device)
        check(L.srg_generator_forward(eng.handle, c_void_p(x.data_ptr()), c_void_p(sr.data_ptr()),
                                      1 if self.training else 0, 1 if self.training else 0, stream_ptr()),
              "srg_generator_forward")
        return sr

    def _eval_into(self, x: torch.Tensor, out: torch.Tensor, packed: set) -> None:
        """eval-mode forward of a chunk of frames straight into ``out`` (a contiguous slice of the caller's batch)."""
        assert x.is_contiguous() and out.is_contiguous()
        n, _, H, W = x.shape
        rt = self._rt
        eng = self._engine(n, H, W, False, x.device, False)
        eng.bind(rt["flat"], None, rt["flat_buf"])
        L = _lib.lib()
        if id(eng) not in packed:
            check(L.srg_generator_pack(eng.handle, stream_ptr()), "srg_generator_pack")
            packed.add(id(eng))
        rt["last_engine"] = eng
        check(L.srg_generator_forward(eng.handle, c_void_p(x.data_ptr()), c_void_p(out.data_ptr()), 0, 0, stream_ptr()),
              "srg_generator_forward")


# ----------------------------------------------------------------------------------------------------------------
# Discriminator
# ----------------------------------------------------------------------------------------------------------------
class _DiscriminatorEngine:
    """One srg_discriminator_t bound to a workspace: fixed (N, H, W) geometry, one in-flight forward at a time."""

    def __init__(self, N: int, H: int, W: int, training: bool, device):
        L = _lib.lib()
        self.handle = c_void_p()
        rc = L.srg_discriminator_create(byref(self.handle), N, H, W)
        if rc != 0:
            msg = L.srg_last_error()
            # same failure mode as the reference (a RuntimeError out of MaxPool2d / InstanceNorm2d)
            raise RuntimeError(msg.decode() if msg else f"srg_discriminator_create failed ({rc})")
        self.N, self.H, self.W, self.training, self.device = N, H, W, training, device
        self.busy = False
        self.ws = None
        self.bound_key = None
        self.param_elems = int(L.srg_discriminator_param_elems(self.handle))
        oh, ow = c_int(), c_int()
        L.srg_discriminator_output_hw(self.handle, byref(oh), byref(ow))
        self.out_hw = (oh.value, ow.value)

    def __del__(self):
        try:
            if self.handle:
                _lib.lib().srg_discriminator_destroy(self.handle)
                self.handle = None
        except Exception:
            pass

    def param_table(self):
        L = _lib.lib()
        out = []
        name = create_string_buffer(128)
        off, numel, ndim = c_int64(), c_int64(), c_int()
        shape = (c_int * 4)()
        for i in range(L.srg_discriminator_num_params(self.handle)):
            check(L.srg_discriminator_param_info(self.handle, i, name, 128, byref(off), byref(numel), byref(ndim), shape))
            out.append((name.value.decode(), off.value, numel.value, tuple(shape[k] for k in range(ndim.value))))
        return out

    def tensor_table(self):
        L = _lib.lib()
        out = {}
        name = create_string_buffer(128)
        off, dt = c_int64(), c_int()
        dims = (c_int * 4)()
        for i in range(L.srg_discriminator_num_tensors(self.handle)):
            check(L.srg_discriminator_tensor_info(self.handle, i, name, 128, byref(off), dims, byref(dt)))
            out[name.value.decode()] = (off.value, tuple(dims[k] for k in range(4)), dt.value)
        return out

    def bind(self, flat_params: torch.Tensor, flat_grads: Optional[torch.Tensor]):
        L = _lib.lib()
        key = flat_params.data_ptr()
        if self.ws is None:
            nbytes = int(L.srg_discriminator_workspace_bytes(self.handle, 1 if self.training else 0))
            self.ws = torch.empty(nbytes + 1024, dtype=torch.uint8, device=self.device)
        if key != self.bound_key:
            base = self.ws.data_ptr()
            aligned = (base + 1023) & ~1023
            check(L.srg_discriminator_bind(self.handle, c_void_p(flat_params.data_ptr()),
                                           c_void_p(flat_grads.data_ptr()) if flat_grads is not None else None,
                                           c_void_p(aligned), self.ws.numel() - (aligned - base), 1 if self.training else 0),
                  "srg_discriminator_bind")
            self.ws_base = aligned
            self.bound_key = key

    def named_tensor(self, name: str) -> torch.Tensor:
        off, dims, dt = self.tensor_table()[name]
        n = dims[0] * dims[1] * dims[2] * dims[3]
        start = (self.ws_base - self.ws.data_ptr()) + off
        if dt == 0:
            return self.ws[start:start + 2 * n].view(torch.bfloat16).view(*dims)
        return self.ws[start:start + 4 * n].view(torch.float32).view(*dims)


class _DiscriminatorFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, module, eng, x, *params):
        L = _lib.lib()
        out = torch.empty(eng.N, 512, eng.out_hw[0], eng.out_hw[1], dtype=torch.float32, device=x.device)
        check(L.srg_discriminator_forward(eng.handle, c_void_p(x.data_ptr()), c_void_p(out.data_ptr()), stream_ptr()),
              "srg_discriminator_forward")
        ctx.module_ref = weakref.ref(module)
        ctx.eng = eng
        _release_on_death(ctx, eng)
        ctx.n_params = len(params)
        ctx.x_shape = tuple(x.shape)
        ctx.param_grads = module._rt.get("param_grads", True)
        ctx.set_materialize_grads(False)
        return out

    @staticmethod
    def backward(ctx, dout):
        module = ctx.module_ref()
        eng = ctx.eng
        if dout is None or module is None:
            _release_if(eng, ctx.busy_token)
            return (None,) * (3 + ctx.n_params)
        L = _lib.lib()
        dout = dout.contiguous()
        if dout.dtype != torch.float32:
            dout = dout.float()
        want_dx = ctx.needs_input_grad[2]
        want_pg = ctx.param_grads and any(ctx.needs_input_grad[3:])
        dx = torch.empty(ctx.x_shape, dtype=torch.float32, device=dout.device) if want_dx else None
        flat_g = None
        if want_pg:
            flat_g = module._grad_buffer_for_backward(eng)
            check(L.srg_discriminator_set_grads(eng.handle, c_void_p(flat_g.data_ptr())))
        check(L.srg_discriminator_backward(eng.handle, c_void_p(dout.data_ptr()), 1 if want_pg else 0,
                                           c_void_p(dx.data_ptr()) if dx is not None else None, stream_ptr()),
              "srg_discriminator_backward")
        _release_if(eng, ctx.busy_token)
        if want_pg:
            module._after_backward(flat_g)
            grads = tuple(flat_g[off:off + n].view(shape) for (_, off, n, shape) in module._ptable)
        else:
            grads = (None,) * ctx.n_params
        return (None, None, dx) + grads


class Discriminator(_FlatModule):
    """Drop-in for the reference's Discriminator (src/models.py:90-120): ``self.model`` is an nn.Sequential whose
    indices 0, 4, 8, 12 hold the conv parameters (state_dict keys ``model.0.weight`` ...).  Input N x 3 x H x W fp32
    CUDA -> N x 512 x h x w sigmoid map.  Raises RuntimeError for inputs the reference network cannot process (every
    axis must be >= 428 and one >= 684)."""

    def __init__(self, input_channels: int = 3, num_filters: int = 64):
        super().__init__()
        if input_channels != 3 or num_filters != 64:
            raise NotImplementedError("libsrgan_b200 implements the reference configuration input_channels=3, num_filters=64")
        nf = num_filters
        layers: List[nn.Module] = []
        specs = [(input_channels, nf, 8, 2), (nf, nf * 2, 4, 1), (nf * 2, nf * 4, 4, 1), (nf * 4, nf * 8, 4, 1)]
        for i, (cin, cout, k, pad) in enumerate(specs):
            layers += [ConvParams(cin, cout, k, stride=2, padding=pad), _Stateless("MaxPool2d(kernel_size=3, stride=2)"),
                       _Stateless(f"InstanceNorm2d({cout})")]
            layers.append(_Stateless("LeakyReLU(0.2)") if i < 3 else _Stateless("Sigmoid()"))
        self.model = nn.Sequential(*layers)
        self._reset_runtime()

    def _tables(self, device):
        rt = self._rt
        if rt.get("probe") is None:
            rt["probe"] = _DiscriminatorEngine(1, 428, 684, False, device)
        probe = rt["probe"]
        return probe.param_table(), [], probe.param_elems, 0

    def _engine(self, N, H, W, training, device):
        pool = self._rt["engines"].setdefault((N, H, W, training), [])
        for e in pool:
            if not e.busy:
                return e
        eng = _DiscriminatorEngine(N, H, W, training, device)
        pool.append(eng)
        return eng

    def last_engine(self):
        return self._rt["last_engine"]

    class _InputGradOnly:
        def __init__(self, module):
            self.m = module

        def __enter__(self):
            self.prev = self.m._rt.get("param_grads", True)
            self.m._rt["param_grads"] = False

        def __exit__(self, *a):
            self.m._rt["param_grads"] = self.prev

    def input_grad_only(self):
        """Context manager: forwards recorded inside back-propagate to the INPUT only (the generator's adversarial
        term, src/train.py:184-190, needs d(D(sr))/d(sr); the reference also fills D's .grad there and then discards
        it with the next d_optimizer.zero_grad())."""
        return Discriminator._InputGradOnly(self)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        if not x.is_cuda:
            raise RuntimeError("Discriminator (libsrgan_b200) runs on a CUDA device only; there is no CPU path")
        if x.dim() != 4 or x.shape[1] != 3:
            raise RuntimeError(f"expected an N x 3 x H x W input, got {tuple(x.shape)}")
        xc = x.contiguous()
        if xc.dtype != torch.float32:
            xc = xc.float()
        N, _, H, W = xc.shape
        self._flatten(xc.device)
        rt = self._rt
        need_grad = torch.is_grad_enabled() and (xc.requires_grad or any(p.requires_grad for p in rt["plist"]))
        eng = self._engine(N, H, W, need_grad, xc.device)
        if need_grad and getattr(eng, "grad_flat", None) is None:
            eng.grad_flat = torch.empty_like(rt["flat"])
        eng.bind(rt["flat"], eng.grad_flat if need_grad else None)
        L = _lib.lib()
        check(L.srg_discriminator_pack(eng.handle, stream_ptr()), "srg_discriminator_pack")
        rt["last_engine"] = eng
        if need_grad:
            _acquire(eng)
            return _DiscriminatorFn.apply(self, eng, xc, *rt["plist"])
        out = torch.empty(N, 512, eng.out_hw[0], eng.out_hw[1], dtype=torch.float32, device=xc.device)
        check(L.srg_discriminator_forward(eng.handle, c_void_p(xc.data_ptr()), c_void_p(out.data_ptr()), stream_ptr()),
              "srg_discriminator_forward")
        return out
