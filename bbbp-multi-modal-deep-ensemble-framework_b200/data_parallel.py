"""Data-parallel training of the reference loop body: replicas + one averaged gradient buffer.

With R ranks at local batch B this equals the reference run (20250113.py:186-191) at batch B with the gradients of R
micro-batches averaged -- NOT the reference at batch R*B, because the encoder's attention scope and BatchNorm's
statistics are per batch.  Reported, not forced: B3DB has ~1 k molecules.

``FlatGradients`` owns ONE persistent fp32 buffer; every ``p.grad`` is a view into it, so autograd accumulates straight
into the buffer (no ``torch.cat`` / copy-back passes over the 54 MB of the MACCS network), and the buffer is averaged
bucket by bucket on a communication stream WHILE the rest of backward is still running: a bucket's all-reduce is enqueued
as soon as the last of its gradients has been accumulated (``register_post_accumulate_grad_hook``).  Buckets follow the
reverse parameter order, i.e. the order backward produces gradients in: head first, encoder / conv weights last.
NCCL averages inside the reduction (``ReduceOp.AVG``); gloo (CPU tests) sums and divides.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


class FlatGradients:
    def __init__(self, parameters, bucket_bytes: int = 16 << 20, group=None):
        self.params = [p for p in parameters if p.requires_grad]
        if not self.params:
            raise ValueError("no trainable parameters")
        dev = self.params[0].device
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
        total = sum(p.numel() for p in self.params)
        self.flat = torch.zeros(total, device=dev, dtype=torch.float32)
        # lay the buffer out in REVERSE parameter order so that a bucket is a contiguous slice completed early in backward
        order = list(reversed(self.params))
        self.buckets, self._bucket_of = [], {}
        off, start, pending = 0, 0, []
        for p in order:
            n = p.numel()
            p.grad = self.flat[off:off + n].view_as(p)
            pending.append(p)
            off += n
            if (off - start) * 4 >= bucket_bytes:
                self._close_bucket(start, off, pending)
                start, pending = off, []
        if pending:
            self._close_bucket(start, off, pending)
        self._left = [len(b["params"]) for b in self.buckets]
        self._works = []
        self.comm_stream = torch.cuda.Stream(dev) if dev.type == "cuda" else None
        self._hooks = [p.register_post_accumulate_grad_hook(self._on_grad) for p in self.params] if self.world > 1 else []

    def _close_bucket(self, start, stop, params):
        idx = len(self.buckets)
        self.buckets.append({"slice": self.flat[start:stop], "params": list(params)})
        for p in params:
            self._bucket_of[p] = idx

    # -- per step ---------------------------------------------------------------------------------------------------------
    def zero(self):
        """Replaces ``optimizer.zero_grad()``: one fill of the flat buffer; the ``p.grad`` views stay in place."""
        if self.flat.is_cuda:
            from . import ops
            ops.fill_zero(self.flat)
        else:
            self.flat.zero_()
        self._left = [len(b["params"]) for b in self.buckets]
        self._works = []

    def _on_grad(self, p):
        b = self._bucket_of[p]
        self._left[b] -= 1
        if self._left[b] == 0:
            self._reduce(b)

    def _reduce(self, b):
        buf = self.buckets[b]["slice"]
        nccl = dist.get_backend(self.group) == "nccl"
        if self.comm_stream is not None:
            self.comm_stream.wait_stream(torch.cuda.current_stream())       # the bucket's gradients are complete
            with torch.cuda.stream(self.comm_stream):
                dist.all_reduce(buf, op=dist.ReduceOp.AVG if nccl else dist.ReduceOp.SUM, group=self.group)
        else:
            dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=self.group)
            buf.div_(self.world)

    def synchronize(self):
        """Call after ``loss.backward()`` and before ``optimizer.step()``: every bucket has been averaged."""
        if self.world == 1:
            return
        for b, left in enumerate(self._left):
            if left > 0:                       # parameters that received no gradient this step: reduce what is there
                self._left[b] = 0
                self._reduce(b)
        if self.comm_stream is not None:
            torch.cuda.current_stream().wait_stream(self.comm_stream)

    def close(self):
        for h in self._hooks:
            h.remove()
        self._hooks = []
