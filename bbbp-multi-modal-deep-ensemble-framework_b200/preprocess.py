"""Device-side input contracts and the PCA projection of the MLP family.

  * ``unpack_zscore``   packed little-endian fingerprint bits -> per-molecule z-scored float32 rows
                        (Descriptors/multi_input_data_preprocess_maccs_opt.py:35-44,121-124)
  * ``u8_image_zscore`` uint8 CHW depiction -> ToTensor scaling -> per-molecule z-score (same file :52-67)
  * ``standardize_chunks`` per-feature z-score inside blocks of 100 molecules
                        (Descriptors/multi_input_data_preprocess_maccs_opt_IsolationForest_fixed_1.py:86-101)
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


def standardize_chunks(features: torch.Tensor, chunk_rows: int = 100) -> torch.Tensor:
    """``StandardScaler().fit_transform`` on every block of ``chunk_rows`` molecules, per FEATURE (column): the
    normalisation behind the ``lso_fixed_1`` pickle the canonical script reads
    (Descriptors/multi_input_data_preprocess_maccs_opt_IsolationForest_fixed_1.py:86-101; 20250113.py:122-126).  The
    reference standardises hstack([MACCS, pixels]); columns are independent, so the two matrices can be passed separately."""
    return ops.standardize_chunks(features if features.stride(1) == 1 else features.contiguous(), chunk_rows)


def pca_transform(x: torch.Tensor, mean: torch.Tensor, components: torch.Tensor, precision: str = "strict") -> torch.Tensor:
    """sklearn ``PCA.transform``: (N, D) float32 -> (N, k) = (x - mean) @ components^T.

    ``precision="strict"`` (default) runs on the tensor cores: the centring is fused into the fp32 -> (hi, lo) fp16 split of
    the rows, the components are split the same way, and one tcgen05 GEMM accumulates hi*hi + lo*hi + hi*lo in fp32 (K =
    49 152 for the depiction features: split-K over eight CTAs per tile).  ``"fp16"`` / ``"bf16"``: one pass.  ``"fp32"``: the
    CUDA-core GEMM (x @ C^T - C @ mean, centring folded into a bias)."""
    x = x if x.is_contiguous() else x.contiguous()
    comp = components if components.is_contiguous() else components.contiguous()
    if precision != "fp32":
        from .autograd import TENSOR_CORE
        fmt, split = TENSOR_CORE[precision]
        K, k = x.shape[1], comp.shape[0]
        a_hi, a_lo = ops.cast16(x, fmt, want_lo=split, sub=mean.contiguous())
        w_hi, w_lo = ops.cast16(comp, fmt, want_lo=split)
        y, _ = ops.gemm_bf16(a_hi, K, w_hi, k, fmt=fmt, a_lo=a_lo, w_lo=w_lo,
                             split_k=ops.fixed_split_k_strict(K) if split else ops.fixed_split_k(K))
        return y
    shift = ops.gemm_f32(mean.reshape(1, -1).contiguous(), comp, trans_b=True)          # (1, k) = mean @ C^T
    neg = ops.scale_by_device_scalar(shift.reshape(-1), torch.full((1,), -1.0, device=x.device))
    return ops.gemm_f32(x, comp, trans_b=True, bias=neg, split_k=ops.fixed_split_k_f32(x.shape[1]))
