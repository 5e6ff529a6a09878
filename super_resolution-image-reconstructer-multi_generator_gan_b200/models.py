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
import math
import weakref
from ctypes import byref, c_char_p, c_int, c_int64, c_void_p, create_string_buffer
from typing import Dict, List, Optional, Tuple

import torch
import torch.nn as nn

from . import _lib
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
        ctx.n_params = len(params)
        ctx.set_materialize_grads(False)
        return sr

    @staticmethod
    def backward(ctx, dsr):
        module = ctx.module_ref()
        eng = ctx.eng
        if dsr is None or module is None:
            eng.busy = False
            return (None,) * (3 + ctx.n_params)
        L = _lib.lib()
        dsr = dsr.contiguous()
        if dsr.dtype != torch.float32:
            dsr = dsr.float()
        flat_g = module._grad_buffer_for_backward()
        check(L.srg_generator_set_grads(eng.handle, c_void_p(flat_g.data_ptr())))
        check(L.srg_generator_backward(eng.handle, c_void_p(dsr.data_ptr()), stream_ptr()), "srg_generator_backward")
        eng.busy = False
        module._after_backward(flat_g)
        grads = tuple(flat_g[off:off + n].view(shape) for (_, off, n, shape) in module._ptable)
        return (None, None, None) + grads


class SRResNet(nn.Module):
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

    # ---- runtime state (never pickled / deep-copied) -----------------------------------------------------------
    def _reset_runtime(self):
        object.__setattr__(self, "_rt", {"flat": None, "flat_buf": None, "nbt": None, "engines": {}, "grad_flat": None,
                                          "grad_hook": None, "sync_bn": False, "last_engine": None})

    def __getstate__(self):
        st = self.__dict__.copy()
        st.pop("_rt", None)
        return st

    def __setstate__(self, st):
        self.__dict__.update(st)
        self._reset_runtime()

    def __deepcopy__(self, memo):
        import copy
        cls = self.__class__
        new = cls.__new__(cls)
        memo[id(self)] = new
        for k, v in self.__dict__.items():
            if k == "_rt":
                continue
            object.__setattr__(new, k, copy.deepcopy(v, memo))
        new._reset_runtime()
        return new

    def _apply(self, fn, recurse=True):
        out = super()._apply(fn, recurse)
        self._rt["flat"] = None           # parameters were re-materialised: re-flatten lazily
        return out

    # ---- flat parameter storage ----------------------------------------------------------------------------------
    def _probe_engine(self, device) -> _GeneratorEngine:
        rt = self._rt
        if rt.get("probe") is None:
            rt["probe"] = _GeneratorEngine(1, 8, 8, self.num_residuals, self.num_upsample_stages, False, device)
        return rt["probe"]

    def _flatten(self, device: torch.device):
        rt = self._rt
        if rt["flat"] is not None and rt["flat"].device == device:
            base = rt["flat"].data_ptr()
            ok = True
            for p, (name, off, n, shape) in zip(rt["plist"], self._ptable):
                if p.data_ptr() != base + 4 * off:
                    ok = False
                    break
            if ok:
                return
        named = dict(self.named_parameters())
        named_buf = dict(self.named_buffers())
        probe = self._probe_engine(device)
        ptable = probe.param_table()
        btable = probe.buffer_table()
        if [t[0] for t in ptable] != list(named.keys()):
            raise RuntimeError("engine parameter table does not match the module's parameters() order")
        with torch.no_grad():
            flat = torch.zeros(probe.param_elems, dtype=torch.float32, device=device)
            for name, off, n, shape in ptable:
                p = named[name]
                if tuple(p.shape) != shape:
                    raise RuntimeError(f"parameter {name} has shape {tuple(p.shape)}, engine expects {shape}")
                view = flat[off:off + n].view(shape)
                view.copy_(p.data)
                p.data = view
            fbuf = torch.zeros(max(probe.buffer_elems, 1), dtype=torch.float32, device=device)
            for name, off, n in btable:
                view = fbuf[off:off + n]
                view.copy_(named_buf[name])
                self._set_buffer(name, view)
            nbt_names = [k for k in named_buf if k.endswith("num_batches_tracked")]
            nbt = torch.zeros(max(len(nbt_names), 1), dtype=torch.long, device=device)
            for i, name in enumerate(nbt_names):
                nbt[i] = named_buf[name].to(device)
                self._set_buffer(name, nbt[i])
        rt["flat"], rt["flat_buf"], rt["nbt"] = flat, fbuf, nbt
        rt["grad_flat"] = None
        rt["grad_store"] = None
        rt["plist"] = [named[t[0]] for t in ptable]
        self._ptable = ptable
        for p in rt["plist"]:
            p._srg_owner = weakref.ref(self)
        for engs in rt["engines"].values():
            for e in engs:
                e.bound_key = None

    def _set_buffer(self, dotted: str, tensor: torch.Tensor):
        mod = self
        parts = dotted.split(".")
        for part in parts[:-1]:
            mod = getattr(mod, part)
        mod._buffers[parts[-1]] = tensor

    def flat_parameters(self) -> torch.Tensor:
        """The flat fp32 buffer every parameter is a view of (engine layout)."""
        dev = next(self.parameters()).device
        self._flatten(dev)
        return self._rt["flat"]

    def flat_grads(self) -> Optional[torch.Tensor]:
        """Flat gradient buffer written by the last backward (same layout as flat_parameters), or None."""
        return self._rt["grad_flat"]

    def _grad_buffer_for_backward(self) -> torch.Tensor:
        rt = self._rt
        flat = rt["flat"]
        g = rt.get("grad_store")
        if g is None or g.device != flat.device or g.numel() != flat.numel():
            g = torch.empty_like(flat)
            rt["grad_store"] = g
        else:
            # gradient accumulation (a second backward before zero_grad): never alias live .grad views
            base, end = g.data_ptr(), g.data_ptr() + 4 * g.numel()
            for p in rt["plist"]:
                if p.grad is not None and base <= p.grad.data_ptr() < end:
                    g = torch.empty_like(flat)
                    break
        return g

    def _after_backward(self, flat_g: torch.Tensor):
        rt = self._rt
        rt["grad_flat"] = flat_g
        hook = rt["grad_hook"]
        if hook is not None:
            hook(self, flat_g)

    def set_grad_hook(self, hook):
        """hook(module, flat_grads) runs right after the engine's backward enqueued its kernels (data-parallel
        gradient all-reduce is installed here, see parallel.py)."""
        self._rt["grad_hook"] = hook

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
                check(_lib.lib().srg_generator_use_nccl(eng.handle), "srg_generator_use_nccl")
            if getattr(self, "debug_keep_grads", False):
                check(_lib.lib().srg_generator_set_keep_grads(eng.handle, 1))
            if rt.get("profile"):
                check(_lib.lib().srg_generator_profile_enable(eng.handle, 1))
            pool.append(eng)
        return eng

    def enable_sync_batchnorm(self):
        """SyncBatchNorm over the communicator created by parallel.init_nccl(): per-channel sums are all-reduced
        between the local reduction and the BatchNorm finalize, forward and backward."""
        self._rt["sync_bn"] = True
        for pool in self._rt["engines"].values():
            for e in pool:
                check(_lib.lib().srg_generator_use_nccl(e.handle), "srg_generator_use_nccl")

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
        eng = self._engine(N, H, W, self.training, x.device, need_grad)
        if self.training and rt.get("grad_store") is None:
            rt["grad_store"] = torch.empty_like(rt["flat"])
        eng.bind(rt["flat"], rt["grad_store"] if self.training else None, rt["flat_buf"])
        L = _lib.lib()
        check(L.srg_generator_pack(eng.handle, stream_ptr()), "srg_generator_pack")
        rt["last_engine"] = eng
        if self.training and rt["nbt"] is not None and self.num_residuals > 0:
            rt["nbt"] += 1
        if need_grad:
            eng.busy = True
            return _GeneratorFn.apply(self, eng, x, *rt["plist"])
        # eval mode (running statistics) or no_grad: no autograd graph.  The reference keeps a graph through an
        # eval-mode generator in train_discriminator (src/train.py:212) but discards those gradients.
        sr = torch.empty(N, 3, H << self.num_upsample_stages, W << self.num_upsample_stages, dtype=torch.float32,
                         device=x.device)
        check(L.srg_generator_forward(eng.handle, c_void_p(x.data_ptr()), c_void_p(sr.data_ptr()),
                                      1 if self.training else 0, 1 if self.training else 0, stream_ptr()),
              "srg_generator_forward")
        return sr
