"""One training step of the reference loop as a single CUDA-graph replay.

The reference trains with (20250113.py:186-191)::

    optimizer.zero_grad(); outputs = model(fingerprints, images).squeeze()
    loss = criterion(outputs, labels); loss.backward(); optimizer.step()

at batch 32 on ~1 k molecules: every kernel of that step is microseconds long, so an eager step is bound by ~340 launches
issued from Python, not by the GPU.  ``GraphedTrainStep`` captures exactly that sequence -- the same autograd graph, the
same bbbp_* kernels, the fused AdamW -- once per input shape and replays it.  What changes between steps travels through
device memory instead of being baked into the capture:

  * AdamW's step count / learning rate (torch LR schedulers keep working) -> bbbp_adamw_dev_f32 + bbbp_adamw_hyper,
  * the dropout seed -> the ``seed_dev`` argument of the dropout / attention kernels,
  * BatchNorm's running statistics and ``num_batches_tracked`` are updated in place by the captured kernels.

The conv branch and the fingerprint encoder are independent until the fusion block; during capture the image branch is
forked onto a second stream so the graph keeps them as parallel branches (forward and backward).
"""
from __future__ import annotations

import os
import struct

import torch

from . import autograd as ag
from . import ops
from .optim import AdamW


class GraphedTrainStep:
    """``step = GraphedTrainStep(model, optimizer, criterion)``; ``loss = step(fingerprint, image, target)``.

    ``model`` is a bbbp_b200 module in train mode, ``optimizer`` a ``bbbp_b200.AdamW`` over its parameters and
    ``criterion`` a ``bbbp_b200.MSELoss`` / ``BCEWithLogitsLoss``.  The returned loss is a device scalar (a copy, safe to
    keep); ``p.grad`` of every parameter holds the step's gradient afterwards, as after an eager step.
    """

    def __init__(self, model, optimizer: AdamW, criterion, squeeze: bool = True, max_graphs: int = 4,
                 fork_image_branch: bool = True):
        if not isinstance(optimizer, AdamW):
            raise TypeError("GraphedTrainStep needs bbbp_b200.AdamW (the captured update reads its scalars from device memory)")
        self.model, self.optimizer, self.criterion = model, optimizer, criterion
        self.squeeze, self.max_graphs, self.fork = squeeze, max_graphs, fork_image_branch
        self._graphs: dict = {}
        self._seed_step = 0
        self._last = None
        # diagnostic switches for the three graph-level optimisations (all on by default; the bit-identity test holds for
        # every combination): conv branch on its own stream, dW / db products on their own stream, chain at high priority
        self.fork_image = True
        self.fork_weight_grads = os.environ.get("BBBP_GRAPH_WGRAD_FORK", "1") == "1"
        self.high_priority_chain = os.environ.get("BBBP_GRAPH_PRIORITY", "1") == "1"

    # -- the step body, shared by warm-up, capture and the eager fallback -------------------------------------------------
    def _body(self, fp, img, y):
        out = self.model(fp, img)
        if self.squeeze:
            out = out.squeeze()
        loss = self.criterion(out, y)
        loss.backward()
        return loss

    def _snapshot(self):
        """Everything a real step mutates, so the warm-up steps needed before capture leave no trace."""
        tensors = [p for p in self.model.parameters()] + [b for b in self.model.buffers()]
        saved = [t.detach().clone() for t in tensors]
        opt_state = {p: {k: (v.clone() if torch.is_tensor(v) else v) for k, v in st.items()}
                     for p, st in self.optimizer.state.items()}
        steps = [g.get("step", 0) for g in self.optimizer.param_groups]
        return tensors, saved, opt_state, steps

    def _restore(self, snap):
        tensors, saved, opt_state, steps = snap
        with torch.no_grad():
            for t, s in zip(tensors, saved):
                t.copy_(s)
            for p, st in self.optimizer.state.items():
                old = opt_state.get(p)
                for k, v in st.items():
                    if torch.is_tensor(v):
                        if old is not None and k in old:
                            v.copy_(old[k])
                        else:
                            v.zero_()
        for g, s in zip(self.optimizer.param_groups, steps):
            g["step"] = s
        ag.clear_weight_cache()

    def _capture(self, fp, img, y):
        dev = fp.device
        model, opt = self.model, self.optimizer
        s_fp, s_img, s_y = torch.empty_like(fp), torch.empty_like(img), torch.empty_like(y)
        s_fp.copy_(fp), s_img.copy_(img), s_y.copy_(y)
        seed_dev = torch.zeros(1, dtype=torch.int64, device=dev)
        plan = opt.graph_prepare()
        snap = self._snapshot()
        prev_seed = ag.set_seed_tensor(seed_dev)
        prev_fork = getattr(model, "fork_image_branch", False)
        try:
            cur = torch.cuda.current_stream(dev)
            side = torch.cuda.Stream(dev)
            side.wait_stream(cur)
            with torch.cuda.stream(side):     # warm-up on a side stream: lazy kernel attributes, TMA maps, allocator pools
                for _ in range(2):
                    opt.zero_grad(set_to_none=True)
                    self._body(s_fp, s_img, s_y)
                    opt.graph_bind(plan)          # eager gradients: exercises the very kernel the capture will hold
                    opt.graph_advance(plan)
                    opt.graph_step(plan)
            cur.wait_stream(side)
            torch.cuda.synchronize(dev)
            self._restore(snap)
            opt.zero_grad(set_to_none=True)
            try:
                graph, s_loss = self._capture_once(s_fp, s_img, s_y, plan, self.fork)
            except Exception:
                if not self.fork:
                    raise
                opt.zero_grad(set_to_none=True)      # multi-stream capture refused: same kernels on one stream
                graph, s_loss = self._capture_once(s_fp, s_img, s_y, plan, False)
            opt.graph_bind(plan)
            grads = [(p, p.grad) for e in plan if e for p in e["params"]]
        finally:
            ag.set_seed_tensor(prev_seed)
            if hasattr(model, "fork_image_branch"):
                model.fork_image_branch = prev_fork
        return dict(graph=graph, fp=s_fp, img=s_img, y=s_y, loss=s_loss, seed=seed_dev, plan=plan, grads=grads)

    def _capture_once(self, s_fp, s_img, s_y, plan, forked):
        graph = torch.cuda.CUDAGraph()
        if hasattr(self.model, "fork_image_branch"):
            self.model.fork_image_branch = forked and self.fork_image
        wgrad = torch.cuda.Stream(s_fp.device) if (forked and self.fork_weight_grads) else None
        prev = ag.set_wgrad_stream(wgrad)
        try:
            # The encoder chain is ~200 microsecond-scale dependent kernels, the forked conv branch a few SM-filling ones:
            # capture the chain on a HIGH-priority stream so its CTAs are dispatched ahead of the conv kernels' pending
            # waves instead of queueing behind them (the branch streams keep the default priority).
            main = torch.cuda.Stream(s_fp.device, priority=-1) if (forked and self.high_priority_chain) else None
            with torch.cuda.graph(graph, stream=main):
                s_loss = self._body(s_fp, s_img, s_y)
                if wgrad is not None:
                    torch.cuda.current_stream().wait_stream(wgrad)     # join the dW / db branch before the update
                self.optimizer.graph_step(plan)
        finally:
            ag.set_wgrad_stream(prev)
        return graph, s_loss

    def __call__(self, fingerprint, image, target):
        if not self.model.training:
            raise RuntimeError("GraphedTrainStep: call model.train() first")
        if fingerprint.is_cuda and fingerprint.device.index != torch.cuda.current_device():
            with torch.cuda.device(fingerprint.device):
                return self(fingerprint, image, target)
        # a capture is tied to the parameter storage it updates in place: .to() / .cuda() after a capture re-captures
        key = (tuple(fingerprint.shape), tuple(image.shape), image.dtype, tuple(target.shape),
               getattr(self.model, "precision", None), fingerprint.device.index, next(self.model.parameters()).data_ptr())
        entry = self._graphs.get(key)
        if entry is None:
            if len(self._graphs) >= self.max_graphs:
                self._graphs.pop(next(iter(self._graphs)))
            entry = self._capture(fingerprint, image, target)
            self._graphs[key] = entry
        if self._last is not entry:       # p.grad must name the buffers THIS graph writes (another shape was replayed last)
            for p, g in entry["grads"]:
                p.grad = g
            self._last = entry
        entry["fp"].copy_(fingerprint, non_blocking=True)
        entry["img"].copy_(image, non_blocking=True)
        entry["y"].copy_(target, non_blocking=True)
        self._seed_step += 1
        seed = (torch.initial_seed() * 0x9E3779B97F4A7C15 + self._seed_step * 0xD1B54A32D192ED03) & 0xFFFFFFFFFFFFFFFF
        ops.store_small(struct.pack("Q", seed), entry["seed"])
        self.optimizer.graph_advance(entry["plan"])
        entry["graph"].replay()
        ag.clear_weight_cache()       # the replay rewrote the parameters behind torch's version counters
        return entry["loss"].detach().clone()
