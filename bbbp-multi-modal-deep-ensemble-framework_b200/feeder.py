"""GPU-resident batch feeder: a drop-in for the reference's ``MixedDataset`` + ``DataLoader`` pair.

The reference builds one ``torch.tensor(row)`` per sample and field on the host, collates 32 of them and copies
6.3 MB to the device per batch (20250113.py:31-45, 165-168, 183-186) -- with the network on tensor cores that host
loop becomes the bottleneck.  Here the fold's arrays are uploaded once; a batch is one gather kernel per field.

Batch composition and ORDER are the reference's: ``DataLoader(shuffle=True)`` draws a seed from torch's global CPU
generator at every ``iter()`` and permutes with ``torch.randperm(n, generator=Generator().manual_seed(seed))``
(torch.utils.data.RandomSampler); the feeder makes exactly those calls, so under the same ``torch.manual_seed`` the
training curve sees the same molecules in the same batches (tested against the real DataLoader on CPU).
"""
from __future__ import annotations

import torch

from . import ops


class DeviceBatchFeeder:
    """``for fingerprints, images, labels in feeder:`` yields device tensors, like the reference's train_loader."""

    def __init__(self, fingerprints, images, labels, batch_size: int = 32, shuffle: bool = False, device="cuda",
                 drop_last: bool = False):
        as_f32 = lambda a: torch.as_tensor(a, dtype=torch.float32)          # MixedDataset casts every field to float32
        self.device = torch.device(device)
        self.fingerprints = as_f32(fingerprints).reshape(len(fingerprints), -1).contiguous().to(self.device)
        self.images = as_f32(images).reshape(len(images), -1).contiguous().to(self.device)
        self.labels = as_f32(labels).reshape(len(labels), 1).contiguous().to(self.device)
        self.batch_size, self.shuffle, self.drop_last = int(batch_size), bool(shuffle), bool(drop_last)
        self.n = self.fingerprints.shape[0]

    def __len__(self):
        return self.n // self.batch_size if self.drop_last else -(-self.n // self.batch_size)

    def epoch_order(self) -> torch.Tensor:
        """Index order of one epoch, consuming torch's global RNG exactly as DataLoader(shuffle=True) does."""
        # DataLoader.__iter__ first draws its per-epoch ``_base_seed`` from the global generator (worker seeding; unused
        # with num_workers=0 but still consumed), then the sampler draws the permutation seed on the first batch
        torch.empty((), dtype=torch.int64).random_()                          # _BaseDataLoaderIter.__init__
        if not self.shuffle:
            return torch.arange(self.n, dtype=torch.int64)
        seed = int(torch.empty((), dtype=torch.int64).random_().item())       # RandomSampler.__iter__
        gen = torch.Generator()
        gen.manual_seed(seed)
        return torch.randperm(self.n, generator=gen)

    def __iter__(self):
        order = self.epoch_order().to(self.device)
        for b in range(len(self)):
            idx = order[b * self.batch_size:(b + 1) * self.batch_size].contiguous()
            if self.device.type == "cuda":
                yield (ops.gather_rows(self.fingerprints, idx), ops.gather_rows(self.images, idx),
                       ops.gather_rows(self.labels, idx).reshape(-1))
            else:   # CPU instances exist only so the host logic (ordering) can be tested without a GPU
                yield self.fingerprints[idx], self.images[idx], self.labels[idx].reshape(-1)
