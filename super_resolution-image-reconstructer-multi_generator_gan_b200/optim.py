"""Adam over the flat parameter buffers of libsrgan_b200 modules.

Replaces ``torch.optim.Adam(model.parameters(), lr=...)`` as used at src/train.py:61-62 (default betas (0.9, 0.999),
eps 1e-8, no weight decay, no amsgrad) with ONE fused kernel launch per model (``srg_adam_step``) instead of torch's
foreach op chain.  It is a ``torch.optim.Optimizer`` so LR schedulers (``LinearLR``, src/train.py:70-71) and
``zero_grad()`` work unchanged.  Parameters that are not views of a flat libsrgan_b200 buffer are rejected.
"""
from __future__ import annotations

from ctypes import c_void_p
from typing import Dict, List

import torch

from . import _lib
from ._lib import check, stream_ptr


class Adam(torch.optim.Optimizer):
    def __init__(self, params, lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8, capturable: bool = False):
        if lr < 0 or eps < 0 or not (0 <= betas[0] < 1) or not (0 <= betas[1] < 1):
            raise ValueError("invalid Adam hyper-parameters")
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps))
        # capturable: step count and learning rate are device scalars, so step() can be recorded into a CUDA graph
        # (train.GraphedGeneratorStep); call sync_lr() after a scheduler changed param_groups[...]["lr"]
        self.capturable = capturable
        self._flat_state: Dict[int, dict] = {}   # id(owner module) -> {"m", "v", "step"}
        self.grad_scale = 1.0                     # multiplies gradients first (e.g. 1/world for a summed all-reduce)

    def _owners(self, group) -> List:
        owners, seen = [], set()
        for p in group["params"]:
            ref = getattr(p, "_srg_owner", None)
            owner = ref() if ref is not None else None
            if owner is None:
                raise RuntimeError("optim.Adam only updates parameters of libsrgan_b200 modules (SRResNet / "
                                   "Discriminator) after they have been moved to their CUDA device")
            if id(owner) not in seen:
                seen.add(id(owner))
                owners.append(owner)
        return owners

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        L = _lib.lib()
        for group in self.param_groups:
            # modules flatten lazily on first CUDA use; make sure that happened
            for p in group["params"]:
                if getattr(p, "_srg_owner", None) is None:
                    raise RuntimeError("optim.Adam: run the model once (or call .flat_parameters()) before step()")
            for owner in self._owners(group):
                flat = owner.flat_parameters()
                plist, ptable = owner._rt["plist"], owner._ptable
                if all(p.grad is None for p in plist):
                    continue                      # no backward since zero_grad(): nothing to do
                st = self._flat_state.get(id(owner))
                if st is None or st["m"].numel() != flat.numel() or st["m"].device != flat.device:
                    st = {"m": torch.zeros_like(flat), "v": torch.zeros_like(flat), "step": 0 if st is None else st["step"]}
                    self._flat_state[id(owner)] = st
                gflat = owner.flat_grads()
                if gflat is None:
                    # user-replaced gradients: gather them back into flat layout (one copy per tensor)
                    gflat = torch.zeros_like(flat)
                    for p, (_, off, n, shape) in zip(plist, ptable):
                        if p.grad is not None:
                            gflat[off:off + n].view(shape).copy_(p.grad)
                b1, b2 = group["betas"]
                if self.capturable:
                    if "step_dev" not in st:
                        st["step_dev"] = torch.full((1,), int(st["step"]), dtype=torch.int32, device=flat.device)
                        st["lr_dev"] = torch.full((1,), float(group["lr"]), dtype=torch.float32, device=flat.device)
                        st["lr_host"] = float(group["lr"])
                    elif st["lr_host"] != float(group["lr"]) and not torch.cuda.is_current_stream_capturing():
                        # a scheduler moved the learning rate (LinearLR, src/train.py:70-71): refresh the device scalar.
                        # Under stream capture the value is baked into nothing -- the graph reads lr_dev -- so the
                        # graphed steps refresh it through sync_lr() right before each replay instead.
                        st["lr_dev"].fill_(float(group["lr"]))
                        st["lr_host"] = float(group["lr"])
                    check(L.srg_adam_step_dev(c_void_p(flat.data_ptr()), c_void_p(gflat.data_ptr()),
                                              c_void_p(st["m"].data_ptr()), c_void_p(st["v"].data_ptr()), flat.numel(),
                                              c_void_p(st["lr_dev"].data_ptr()), float(b1), float(b2), float(group["eps"]),
                                              c_void_p(st["step_dev"].data_ptr()), float(self.grad_scale), stream_ptr()),
                          "srg_adam_step_dev")
                    continue
                st["step"] += 1
                check(L.srg_adam_step(c_void_p(flat.data_ptr()), c_void_p(gflat.data_ptr()), c_void_p(st["m"].data_ptr()),
                                      c_void_p(st["v"].data_ptr()), flat.numel(), float(group["lr"]), float(b1), float(b2),
                                      float(group["eps"]), int(st["step"]), float(self.grad_scale), stream_ptr()),
                      "srg_adam_step")
        return loss

    def sync_lr(self) -> None:
        """capturable mode: push param_groups' learning rates to their device scalars (outside graph capture)."""
        for group in self.param_groups:
            for owner in self._owners(group):
                st = self._flat_state.get(id(owner))
                if st is not None and "lr_dev" in st and st["lr_host"] != float(group["lr"]):
                    st["lr_dev"].fill_(float(group["lr"]))
                    st["lr_host"] = float(group["lr"])

    # ---- checkpointing: m / v / step live in flat per-module buffers, not in torch's per-parameter state dict
    def state_dict(self):
        sd = super().state_dict()
        flat = []
        for group in self.param_groups:
            for owner in self._owners(group):
                st = self._flat_state.get(id(owner))
                if st is None:
                    flat.append(None)
                    continue
                step = int(st["step_dev"].item()) if "step_dev" in st else int(st["step"])
                flat.append({"m": st["m"].detach().clone(), "v": st["v"].detach().clone(), "step": step})
        sd["srg_flat_state"] = flat
        return sd

    def load_state_dict(self, state_dict):
        state_dict = dict(state_dict)
        flat = state_dict.pop("srg_flat_state", None)
        super().load_state_dict(state_dict)
        if flat is None:
            return
        i = 0
        for group in self.param_groups:
            for owner in self._owners(group):
                src = flat[i] if i < len(flat) else None
                i += 1
                if src is None:
                    continue
                ref = owner.flat_parameters()
                st = {"m": src["m"].to(ref.device).clone(), "v": src["v"].to(ref.device).clone(), "step": int(src["step"])}
                if st["m"].numel() != ref.numel():
                    raise RuntimeError("optim.Adam.load_state_dict: flat state does not match the module's parameter layout")
                self._flat_state[id(owner)] = st          # capturable device scalars are re-created lazily by step()

    def flat_state(self, owner) -> dict:
        """{"m": exp_avg, "v": exp_avg_sq, "step": int} in the owner's flat layout (checkpointing / tests)."""
        return self._flat_state.get(id(owner))
