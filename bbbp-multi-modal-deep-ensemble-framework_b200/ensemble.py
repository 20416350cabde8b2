"""Ensemble glue: the NN member's column of the stacking design matrix, without a host sync per batch.

The reference collects the network's out-of-fold predictions batch by batch with ``.cpu()`` + ``list.extend`` and
writes them to their dataset positions (20250113.py:227-238: ``nn_predictions[test_idx] = nn_fold_predictions``), then
stacks the members' columns for the meta-learner (20250113.py:394-404: ``np.vstack([nn, rf, xgb, cat]).T``;
_opt_more.py:191-203 with ``Ridge``).  The tree members and the meta-learner stay on the reference's CPU code
(north_star); what moves here is the NN column: a fold is scored on the device with the reference's batch semantics,
scattered to its positions by one kernel, and the whole column crosses PCIe once.
"""
from __future__ import annotations

import numpy as np
import torch

from . import ops


class OutOfFoldScores:
    """Device-resident ``nn_predictions`` vector of a K-fold run."""

    def __init__(self, n_molecules: int, device):
        self.scores = torch.zeros((n_molecules,), device=device, dtype=torch.float32)   # np.zeros(len(y)) in the reference
        self.filled = torch.zeros((n_molecules,), device=device, dtype=torch.float32)

    @torch.no_grad()
    def add_fold(self, model, fingerprint, image, test_idx, batch_size: int = 32):
        """Score the fold's test molecules in ``test_idx`` order (the order a ``DataLoader(shuffle=False)`` over the
        fold's test subset yields, so batch composition -- which the cross-molecule attention makes part of the result --
        matches the reference) and write them to their dataset positions.  ``fingerprint`` / ``image`` are the fold's
        test tensors on the device, row r belonging to molecule ``test_idx[r]``."""
        idx = torch.as_tensor(np.asarray(test_idx), dtype=torch.int64).to(self.scores.device)
        if idx.numel() != fingerprint.shape[0]:
            raise ValueError(f"{idx.numel()} indices for {fingerprint.shape[0]} molecules")
        was_training = model.training
        model.eval()
        try:
            if hasattr(model, "predict_batches"):
                fold = model.predict_batches(fingerprint, image, batch_size)
            else:                                   # MLP family: per-molecule network, batch composition is irrelevant
                fold = model(fingerprint, image).reshape(-1)
        finally:
            model.train(was_training)
        ops.scatter(fold.contiguous(), idx, self.scores)
        ops.scatter(torch.ones_like(fold), idx, self.filled)
        return fold

    def to_numpy(self) -> np.ndarray:
        """The single device->host transfer of the column."""
        return self.scores.cpu().numpy()

    def complete(self) -> bool:
        return bool(self.filled.min().item() == 1.0)


def stack_columns(nn_column, *member_columns) -> np.ndarray:
    """``np.vstack([nn, rf, xgb, cat]).T`` (20250113.py:403): the (N, members) matrix the reference's
    ``StackingRegressor`` / ``Ridge`` / ``LinearRegression`` meta-learners are fitted on.  ``nn_column`` may be an
    OutOfFoldScores, a device tensor or a host array; the other members' columns are host arrays from the CPU models."""
    if isinstance(nn_column, OutOfFoldScores):
        nn_column = nn_column.to_numpy()
    elif torch.is_tensor(nn_column):
        nn_column = nn_column.detach().cpu().numpy()
    cols = [np.asarray(nn_column, dtype=np.float64)] + [np.asarray(c, dtype=np.float64) for c in member_columns]
    n = cols[0].shape[0]
    for c in cols:
        if c.shape != (n,):
            raise ValueError(f"member column of shape {c.shape}, expected ({n},)")
    return np.vstack(cols).T
