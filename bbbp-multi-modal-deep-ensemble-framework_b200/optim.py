"""Fused multi-tensor AdamW: one kernel launch updates every parameter tensor.

Semantics follow ``torch.optim.AdamW`` single-tensor math as the reference configures it
(``optim.AdamW(model.parameters(), lr=1e-4, weight_decay=1e-5)``, 20250113.py:172,191):
decoupled decay, bias-corrected moments, eps added after the sqrt.  ``lr`` is read from the
param group at every step, so torch LR schedulers (``CosineAnnealingWarmRestarts`` in
_transformer_cnn.py:161, ``ReduceLROnPlateau`` in _opt_more.py:162) keep working.
"""
from __future__ import annotations

import struct

import torch

from . import ops

from ._lib import lib as _lib

_CHUNK = _lib.bbbp_adamw_chunk()      # elements per CTA of the fused update


def _zeros_like(p: torch.Tensor) -> torch.Tensor:
    """Moment buffers zeroed by the library's fill kernel (contiguous CUDA parameters; anything else falls back to torch)."""
    if p.is_cuda and p.is_contiguous():
        return ops.fill_zero(torch.empty_like(p))
    return torch.zeros_like(p, memory_format=torch.preserve_format)


class AdamW(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2):
        defaults = dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay)
        super().__init__(params, defaults)
        self._tables = {}
        self._step_mirror = {}

    # -- state layout shared with torch.optim.AdamW ------------------------------------------------------------------------
    # torch keeps the step count per parameter (state[p]["step"], a CPU float32 scalar); this optimizer counts per group
    # (one launch updates the whole group).  Both are maintained so checkpoints interchange in either direction: a stock
    # AdamW state_dict loaded here resumes bias correction at its step, and ours loads into stock AdamW.
    def _group_step(self, gi, group, params) -> int:
        if "step" not in group:
            prior = [float(self.state[p]["step"]) for p in params if "step" in self.state.get(p, {})]
            group["step"] = int(max(prior)) if prior else 0
        return group["step"]

    def _mirror_step(self, gi, group, params) -> None:
        shared = self._step_mirror.get(gi)
        if shared is None:
            shared = self._step_mirror[gi] = torch.zeros((), dtype=torch.float32)
        shared.fill_(float(group["step"]))
        for p in params:
            st = self.state[p]
            if st.get("step") is not shared:
                st["step"] = shared

    def state_dict(self):
        """torch's layout: every parameter gets its OWN ``step`` tensor (the in-memory mirror is one shared scalar per group;
        stock AdamW increments ``state[p]["step"]`` in place per parameter, so a shared object would be bumped N times)."""
        sd = super().state_dict()
        for st in sd["state"].values():
            if torch.is_tensor(st.get("step")):
                st["step"] = st["step"].clone()
        return sd

    @staticmethod
    def _to_device(values, dtype, dev):
        """Small host table -> device through pinned memory (no synchronous pageable copy in the middle of training)."""
        return torch.tensor(values, dtype=dtype, pin_memory=True).to(dev, non_blocking=True)

    def _table(self, gi, params):
        """Device tables for one param group: [param | grad | exp_avg | exp_avg_sq] pointers, sizes, chunks."""
        key = tuple((p.data_ptr(), p.grad.data_ptr()) for p in params)
        hit = self._tables.get(gi)
        if hit is not None and hit[0] == key:
            return hit[1]
        dev = params[0].device
        ptrs, sizes, chunk_t, chunk_o = [], [], [], []
        for p in params:
            st = self.state[p]
            if "exp_avg" not in st:
                st["exp_avg"] = _zeros_like(p)
                st["exp_avg_sq"] = _zeros_like(p)
        for sel in (lambda p: p, lambda p: p.grad, lambda p: self.state[p]["exp_avg"],
                    lambda p: self.state[p]["exp_avg_sq"]):
            ptrs += [sel(p).data_ptr() for p in params]
        for t, p in enumerate(params):
            n = p.numel()
            sizes.append(n)
            for off in range(0, n, _CHUNK):
                chunk_t.append(t)
                chunk_o.append(off)
        tab = (self._to_device(ptrs, torch.int64, dev), self._to_device(sizes, torch.int64, dev),
               self._to_device(chunk_t, torch.int32, dev), self._to_device(chunk_o, torch.int64, dev),
               len(params), len(chunk_t))
        self._tables[gi] = (key, tab)
        return tab

    # -- CUDA-graph form (train.GraphedTrainStep) ---------------------------------------------------------------------
    def graph_prepare(self):
        """Before capture: moment buffers and EMPTY device tables for every group (their contents are only read when
        the captured kernel runs, so they are filled by graph_bind() once the capture has fixed the gradient
        addresses).  Returns one (8-float hyper tensor) per group."""
        plan = []
        for group in self.param_groups:
            params = [p for p in group["params"] if p.requires_grad]
            if not params:
                plan.append(None)
                continue
            dev = params[0].device
            for p in params:
                if not (p.is_cuda and p.dtype == torch.float32 and p.is_contiguous()):
                    raise RuntimeError("bbbp_b200.AdamW needs contiguous float32 CUDA parameters")
                st = self.state[p]
                if "exp_avg" not in st:
                    st["exp_avg"] = _zeros_like(p)
                    st["exp_avg_sq"] = _zeros_like(p)
            sizes, chunk_t, chunk_o = [], [], []
            for t, p in enumerate(params):
                sizes.append(p.numel())
                for off in range(0, p.numel(), _CHUNK):
                    chunk_t.append(t)
                    chunk_o.append(off)
            plan.append(dict(params=params, ptrs=torch.zeros(4 * len(params), dtype=torch.int64, device=dev),
                             sizes=torch.tensor(sizes, dtype=torch.int64).to(dev),
                             chunk_t=torch.tensor(chunk_t, dtype=torch.int32).to(dev),
                             chunk_o=torch.tensor(chunk_o, dtype=torch.int64).to(dev), nchunks=len(chunk_t),
                             hyper=torch.zeros(8, dtype=torch.float32, device=dev)))
        return plan

    def graph_step(self, plan):
        """Inside capture: one adamw_dev launch per group reading tables + hyper-parameters from device memory."""
        for entry in plan:
            if entry is not None:
                ops.adamw_dev(entry["ptrs"], entry["sizes"], entry["chunk_t"], entry["chunk_o"], len(entry["params"]),
                              entry["nchunks"], entry["hyper"])

    def graph_bind(self, plan):
        """After capture: the gradients now have their (static) graph-pool addresses."""
        for entry in plan:
            if entry is None:
                continue
            params = entry["params"]
            missing = [i for i, p in enumerate(params) if p.grad is None]
            if missing:
                raise RuntimeError(f"{len(missing)} parameters received no gradient in the captured step")
            ptrs = []
            for sel in (lambda p: p, lambda p: p.grad, lambda p: self.state[p]["exp_avg"],
                        lambda p: self.state[p]["exp_avg_sq"]):
                ptrs += [sel(p).data_ptr() for p in params]
            entry["ptrs"].copy_(torch.tensor(ptrs, dtype=torch.int64))

    def graph_advance(self, plan, grad_scale: float = 1.0):
        """Before each replay: bump the step counts and send this step's scalars (lr may have been changed by a torch
        LR scheduler) to the device as kernel parameters (bbbp_store_small)."""
        for gi, (group, entry) in enumerate(zip(self.param_groups, plan)):
            if entry is None:
                continue
            group["step"] = self._group_step(gi, group, entry["params"]) + 1
            self._mirror_step(gi, group, entry["params"])
            b1, b2 = group["betas"]
            h = ops.adamw_hyper(group["lr"], b1, b2, group["eps"], group["weight_decay"], group["step"], grad_scale)
            ops.store_small(struct.pack("8f", *h), entry["hyper"])

    @torch.no_grad()
    def step(self, closure=None, grad_scale: float = 1.0):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        for gi, group in enumerate(self.param_groups):
            params = [p for p in group["params"] if p.grad is not None]
            if not params:
                continue
            for p in params:
                if not (p.is_cuda and p.dtype == torch.float32 and p.is_contiguous() and p.grad.is_contiguous()):
                    raise RuntimeError("bbbp_b200.AdamW needs contiguous float32 CUDA parameters and gradients")
            group["step"] = self._group_step(gi, group, params) + 1
            ptrs, sizes, chunk_t, chunk_o, nt, nc = self._table(gi, params)
            self._mirror_step(gi, group, params)
            b1, b2 = group["betas"]
            ops.adamw(ptrs, sizes, chunk_t, chunk_o, nt, nc, group["lr"], b1, b2, group["eps"], group["weight_decay"],
                      group["step"], grad_scale)
        # the kernel wrote the parameters behind torch's version counters: drop derived bf16 weight copies
        from .autograd import clear_weight_cache
        clear_weight_cache()
        return loss
