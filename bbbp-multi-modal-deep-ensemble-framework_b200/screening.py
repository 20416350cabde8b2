"""Batch-sharded virtual screening across 1/2/4/8 GPUs.

The reference scores molecules batch by batch (20250113.py:229-237) and, because the encoder attends
across the molecules of a batch, a reference batch is the atomic unit of work: molecules
[b*Bs, (b+1)*Bs) form batch b no matter how many GPUs take part.  Rank r of R owns a contiguous block
of whole batches; there is no data-path collective, only one gather of the fp32 scores at the end.
Results are therefore bit-identical for every R.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def partition_batches(n_molecules: int, batch_size: int, world_size: int, rank: int) -> tuple[int, int]:
    """Molecule range [start, stop) of ``rank``: batches [floor(r*nb/R), floor((r+1)*nb/R)), where the
    ragged tail batch (n mod batch_size molecules) is the last batch and so lands on the last rank."""
    if n_molecules < 0 or batch_size <= 0 or world_size <= 0 or not 0 <= rank < world_size:
        raise ValueError((n_molecules, batch_size, world_size, rank))
    nb = -(-n_molecules // batch_size)
    b0 = rank * nb // world_size
    b1 = (rank + 1) * nb // world_size
    return min(n_molecules, b0 * batch_size), min(n_molecules, b1 * batch_size)


def gather_scores(local: torch.Tensor, n_molecules: int, batch_size: int, group=None) -> torch.Tensor:
    """All-gather the per-rank score slices into the (n_molecules,) vector (one collective; NCCL over
    NVLink on GPUs, gloo on CPU tensors in the tests).  Slices are padded to the largest count."""
    if not (dist.is_available() and dist.is_initialized()):
        assert local.numel() == n_molecules
        return local
    world = dist.get_world_size(group)
    spans = [partition_batches(n_molecules, batch_size, world, r) for r in range(world)]
    counts = [b - a for a, b in spans]
    width = max(max(counts), 1)
    send = torch.zeros((width,), device=local.device, dtype=local.dtype)
    send[: local.numel()].copy_(local.reshape(-1))
    recv = torch.empty((world * width,), device=local.device, dtype=local.dtype)
    dist.all_gather_into_tensor(recv, send, group=group)
    out = torch.empty((n_molecules,), device=local.device, dtype=local.dtype)
    for r, (a, b) in enumerate(spans):
        out[a:b].copy_(recv[r * width: r * width + (b - a)])
    return out


@torch.no_grad()
def screen(model, load_molecules, n_molecules: int, batch_size: int = 256, chunk_molecules: int = 8192, group=None,
           gather: bool = True) -> torch.Tensor:
    """Score ``n_molecules`` with the reference's batch semantics, sharded by whole batches.

    ``load_molecules(start, stop)`` returns this rank's (fingerprint, image) CUDA float32 tensors for the
    global molecule range [start, stop) -- always whole reference batches except the global tail.
    Returns the full (n_molecules,) score vector on every rank (or the local slice if gather=False)."""
    rank = dist.get_rank(group) if dist.is_available() and dist.is_initialized() else 0
    world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
    a, b = partition_batches(n_molecules, batch_size, world, rank)
    device = next(model.parameters()).device
    local = torch.empty((b - a,), device=device, dtype=torch.float32)
    chunk = max(1, chunk_molecules // batch_size) * batch_size
    for start in range(a, b, chunk):
        stop = min(b, start + chunk)
        fp, img = load_molecules(start, stop)
        local[start - a: stop - a].copy_(model.predict_batches(fp, img, batch_size, max_rows_per_pass=chunk))
    return gather_scores(local, n_molecules, batch_size, group) if gather else local


def average_gradients(parameters, group=None) -> None:
    """Data-parallel training step helper: one all-reduce of the flat fp32 gradient buffer (54 MB for the
    MACCS network) followed by 1/R.  With R ranks at local batch B this equals the reference run at batch B with the
    gradients of R micro-batches averaged -- NOT the reference at batch R*B, because attention scope and BatchNorm
    statistics are per batch.  Reported, not forced: B3DB has ~1 k molecules."""
    if not (dist.is_available() and dist.is_initialized()):
        return
    world = dist.get_world_size(group)
    grads = [p.grad for p in parameters if p.grad is not None]
    if not grads or world == 1:
        return
    flat = torch.cat([g.reshape(-1) for g in grads])         # flat fp32 gradient buffer (layout only)
    if dist.get_backend(group) == "nccl":
        dist.all_reduce(flat, op=dist.ReduceOp.AVG, group=group)     # the 1/R lives inside the NCCL reduction
    else:                                                             # gloo (CPU tests) has no AVG
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
        flat.div_(world)
    off = 0
    for g in grads:
        n = g.numel()
        g.copy_(flat[off:off + n].view_as(g))
        off += n


class SparseDepictions:
    """Lossless sparse encoding of uint8 (N, 3, 128, 128) depictions for host-fed screening (csrc/sparse_depictions.cu): a
    bit mask of the non-white pixels (2 048 bytes per molecule), their RGB triples in scan order and running pixel counts.
    RDKit depictions are ~93 % white, so a molecule takes ~5.5 KB instead of 49 152 bytes over PCIe;
    ``predict_from_host(packed_bits, SparseDepictions, ..., packed=True)`` decodes on the device and scores exactly the
    same uint8 images."""

    def __init__(self, mask: torch.Tensor, values: torch.Tensor, offsets: torch.Tensor):
        self.mask, self.values, self.offsets = mask, values, offsets
        self.shape = (mask.shape[0], 3, 128, 128)
        self.dtype = torch.uint8

    @classmethod
    def encode(cls, images_u8, pin: bool = True) -> "SparseDepictions":
        import numpy as np
        img = np.ascontiguousarray(np.asarray(images_u8, dtype=np.uint8)).reshape(-1, 3, 128 * 128)
        marked = (img != 255).any(axis=1)                                        # (N, 16384)
        mask = np.packbits(marked, axis=1, bitorder="little")                    # (N, 2048)
        values = np.ascontiguousarray(img.transpose(0, 2, 1)[marked])            # (T, 3) in (molecule, pixel) scan order
        offsets = np.zeros(img.shape[0] + 1, dtype=np.int64)
        np.cumsum(marked.sum(axis=1), out=offsets[1:])
        values = values.reshape(-1)
        if values.size == 0:
            values = np.zeros(16, dtype=np.uint8)
        out = [torch.from_numpy(mask), torch.from_numpy(values), torch.from_numpy(offsets)]
        if pin and torch.cuda.is_available():
            out = [t.pin_memory() for t in out]
        return cls(*out)

    def decode_host(self):
        """numpy reconstruction (the oracle of the device kernel)."""
        import numpy as np
        n = self.mask.shape[0]
        marked = np.unpackbits(self.mask.numpy(), axis=1, bitorder="little").astype(bool)
        img = np.full((n, 128 * 128, 3), 255, dtype=np.uint8)
        total = int(self.offsets[-1])
        img[marked] = self.values.numpy()[: 3 * total].reshape(-1, 3)
        return img.transpose(0, 2, 1).reshape(n, 3, 128, 128)

    def nbytes(self) -> int:
        return int(self.mask.numel() + 3 * int(self.offsets[-1]) + self.offsets.numel() * 8)


def pack_fingerprint_bits(bits) -> "torch.Tensor":
    """(N, F) 0/1 array (the rows ``create_descriptors_zinc.py:62`` saves to ``morgan_fingerprints.npy``) -> (N, ceil(F/8))
    uint8, little-endian bit order: the packed input contract of predict_batches_packed (8x to 32x fewer H2D bytes)."""
    import numpy as np
    return torch.from_numpy(np.packbits(np.asarray(bits).astype(np.uint8), axis=1, bitorder="little"))


@torch.no_grad()
def screen_library(model, fingerprints, depictions, out_csv: str | None = None, ids=None, smiles=None, batch_size: int = 256,
                   chunk_molecules: int = 1024, classification: bool = False, threshold: float = 0.5, group=None):
    """File-level screening driver with the shape of ``Descriptors/virtualscreening.py:1-19``: a precomputed library ->
    one score per molecule -> a results table with the reference's ``Prediction`` / ``Probability`` columns.

    ``fingerprints``: path to / array of (N, F) 0/1 bits (``morgan_fingerprints.npy``, create_descriptors_zinc.py:62) or of
    already packed (N, ceil(F/8)) uint8 rows; ``depictions``: path to / array of (N, 3, 128, 128) uint8 images (the 2D
    depiction PNGs of convert_smiles_2_img.py decoded and resized offline; RDKit / PIL stay on the host).  ``.npy`` paths
    are memory-mapped and staged through pinned buffers one shard at a time.  Under torch.distributed every rank scores
    its block of whole reference batches (partition_batches) and rank 0 writes the table.  Regression heads report the
    score as ``Prediction``; ``classification=True`` treats it as a logit: ``Probability = sigmoid(score)``,
    ``Prediction = Probability >= threshold`` (virtualscreening.py:13-14 with the NN in place of rf_model)."""
    import numpy as np
    fp = np.load(fingerprints, mmap_mode="r") if isinstance(fingerprints, str) else np.asarray(fingerprints)
    img = np.load(depictions, mmap_mode="r") if isinstance(depictions, str) else np.asarray(depictions)
    n = fp.shape[0]
    if img.shape[0] != n:
        raise ValueError(f"{n} fingerprints but {img.shape[0]} depictions")
    n_bits = model.fingerprint_transformer.layers[0].self_attn.embed_dim
    packed_in = fp.dtype == np.uint8 and fp.shape[1] == (n_bits + 7) // 8
    if not packed_in and fp.shape[1] != n_bits:
        raise ValueError(f"fingerprint width {fp.shape[1]} matches neither {n_bits} bits nor {(n_bits + 7) // 8} packed bytes")
    rank = dist.get_rank(group) if dist.is_available() and dist.is_initialized() else 0
    world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
    a, b = partition_batches(n, batch_size, world, rank)
    device = next(model.parameters()).device
    local = torch.empty((b - a,), device=device, dtype=torch.float32)
    shard = max(1, 65536 // batch_size) * batch_size                       # molecules staged in pinned memory at a time
    was_training = model.training
    model.eval()
    try:
        for s in range(a, b, shard):
            e = min(b, s + shard)
            rows = np.array(fp[s:e], copy=True)
            packed = (torch.from_numpy(rows) if packed_in else pack_fingerprint_bits(rows)).pin_memory()
            images = torch.from_numpy(np.array(img[s:e], copy=True)).pin_memory()      # mmap slices are read-only
            _, dev_scores = model.predict_from_host(packed, images, batch_size, chunk_molecules=chunk_molecules, packed=True,
                                                    return_device=True)
            local[s - a: e - a].copy_(dev_scores)
    finally:
        model.train(was_training)
    scores = gather_scores(local, n, batch_size, group).cpu().numpy()
    table = {}
    if ids is not None:
        table["ZINC_ID"] = list(ids)
    if smiles is not None:
        table["SMILES"] = list(smiles)
    if classification:
        prob = 1.0 / (1.0 + np.exp(-scores.astype(np.float64)))
        table["Prediction"] = (prob >= threshold).astype(np.int64)
        table["Probability"] = prob
    else:
        table["Prediction"] = scores
    if out_csv is not None and rank == 0:
        import pandas as pd
        pd.DataFrame(table).to_csv(out_csv, index=False)                   # virtualscreening.py:19
    return table
