"""Flat parameter storage shared by the libsrgan_b200 modules (SRResNet, Discriminator).

Every ``nn.Parameter`` of a module is a view into ONE flat fp32 buffer laid out as the engine's parameter table;
gradients come back as views of one flat buffer of the same layout, so the optimiser and the data-parallel gradient
all-reduce are single launches over contiguous memory.  Runtime state (flat buffers, engines, hooks) lives in
``self._rt`` and is never pickled or deep-copied.
"""
from __future__ import annotations

import weakref
from typing import Optional

import torch
import torch.nn as nn


class _FlatModule(nn.Module):
    def _tables(self, device):  # -> (param table, buffer table, param_elems, buffer_elems)
        raise NotImplementedError

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
        ptable, btable, param_elems, buffer_elems = self._tables(device)
        if [t[0] for t in ptable] != list(named.keys()):
            raise RuntimeError("engine parameter table does not match the module's parameters() order")
        with torch.no_grad():
            flat = torch.zeros(param_elems, dtype=torch.float32, device=device)
            for name, off, n, shape in ptable:
                p = named[name]
                if tuple(p.shape) != shape:
                    raise RuntimeError(f"parameter {name} has shape {tuple(p.shape)}, engine expects {shape}")
                view = flat[off:off + n].view(shape)
                view.copy_(p.data)
                p.data = view
            fbuf = torch.zeros(max(buffer_elems, 1), dtype=torch.float32, device=device)
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
        """The flat gradient buffer the live ``.grad`` views alias (same layout as flat_parameters), or None when
        there are no gradients or they do not live in one of this module's flat buffers (e.g. autograd summed the
        contributions of two live forwards into fresh tensors)."""
        rt = self._rt
        plist = rt.get("plist")
        if not plist or plist[0].grad is None:
            return None
        base = plist[0].grad.data_ptr() - 4 * self._ptable[0][1]
        cands = [rt.get("grad_flat")] + [getattr(e, "grad_flat", None) for pool in rt["engines"].values() for e in pool]
        for cand in cands:
            if cand is None or cand.data_ptr() != base:
                continue
            for p, (_, off, n, _) in zip(plist, self._ptable):
                if p.grad is None or p.grad.data_ptr() != base + 4 * off:
                    return None
            return cand
        return None

    def _grad_buffer_for_backward(self, eng) -> torch.Tensor:
        """Every engine (= every live forward) owns one flat gradient buffer, so two forwards awaiting backward never
        share one; a buffer still aliased by live ``.grad`` views (gradient accumulation without zero_grad) is not
        reused either."""
        rt = self._rt
        flat = rt["flat"]
        g = getattr(eng, "grad_flat", None)
        if g is None or g.device != flat.device or g.numel() != flat.numel():
            g = torch.empty_like(flat)
            eng.grad_flat = g
            return g
        base, end = g.data_ptr(), g.data_ptr() + 4 * g.numel()
        for p in rt["plist"]:
            if p.grad is not None and base <= p.grad.data_ptr() < end:
                return torch.empty_like(flat)
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

