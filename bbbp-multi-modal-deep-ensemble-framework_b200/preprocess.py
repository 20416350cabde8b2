"""Device-side input contracts and the PCA projection of the MLP family.

  * ``unpack_zscore``   packed little-endian fingerprint bits -> per-molecule z-scored float32 rows
                        (Descriptors/multi_input_data_preprocess_maccs_opt.py:35-44,121-124)
  * ``u8_image_zscore`` uint8 CHW depiction -> ToTensor scaling -> per-molecule z-score (same file :52-67)
  * ``pca_transform``   sklearn ``PCA.transform`` = (X - mean) @ components.T as used at
                        Models/multi_input_data_regression_opt_transformer_cnn_opt.py:30-33 (``fit`` stays on sklearn)
Oracle for all three: oracle/preprocess.py.
"""
from __future__ import annotations

import torch

from . import ops


def unpack_zscore(packed: torch.Tensor, n_bits: int) -> torch.Tensor:
    return ops.unpack_zscore(packed, n_bits)


def u8_image_zscore(images_u8: torch.Tensor) -> torch.Tensor:
    return ops.u8_zscore(images_u8)


def pca_transform(x: torch.Tensor, mean: torch.Tensor, components: torch.Tensor) -> torch.Tensor:
    """(N, D) float32 -> (N, k): x @ C^T - (C @ mean), two launches of the fp32 GEMM (the centring is folded into a
    bias so the (N, D) centred matrix is never written)."""
    x = x if x.is_contiguous() else x.contiguous()
    comp = components if components.is_contiguous() else components.contiguous()
    shift = ops.gemm_f32(mean.reshape(1, -1).contiguous(), comp, trans_b=True)          # (1, k) = mean @ C^T
    neg = ops.scale_by_device_scalar(shift.reshape(-1), torch.full((1,), -1.0, device=x.device))
    return ops.gemm_f32(x, comp, trans_b=True, bias=neg, split_k=ops.fixed_split_k_f32(x.shape[1]))
