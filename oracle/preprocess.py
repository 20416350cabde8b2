"""numpy restatement of the input contracts P1/P2/P16 (TEST INFRASTRUCTURE, see oracle/__init__.py).

  * P1  per-molecule z-score of fingerprint bits:
        /root/reference/Descriptors/multi_input_data_preprocess_maccs_opt.py:121-124
        ``StandardScaler().fit_transform(bits.reshape(-1, 1))`` -> mean/std over the F bits of
        ONE molecule, population std (ddof 0), float64 math; consumers cast to float32
        (MixedDataset.__getitem__, 20250113.py:40-45).  sklearn maps a zero std to 1.
  * P2  image standardisation, same call on the 49 152 pixel values of one molecule.
  * P16 PCA projection ``(X - mean) @ components.T``: sklearn PCA.transform as used at
        Models/multi_input_data_regression_opt_transformer_cnn_opt.py:30-33 (whiten=False).
  * packed-bit input (extension, no reference code): bit i of a molecule lives in byte
        i // 8, bit position i % 8 (little-endian bit order) == np.unpackbits(bitorder="little").
"""
from __future__ import annotations

import numpy as np


def unpack_bits(packed: np.ndarray, n_bits: int) -> np.ndarray:
    """(B, ceil(F/8)) uint8 -> (B, F) uint8 of 0/1."""
    return np.unpackbits(np.ascontiguousarray(packed, dtype=np.uint8), axis=1, bitorder="little")[:, :n_bits]


def pack_bits(bits: np.ndarray) -> np.ndarray:
    return np.packbits(np.ascontiguousarray(bits, dtype=np.uint8), axis=1, bitorder="little")


def zscore_rows(values: np.ndarray) -> np.ndarray:
    """Per-row standardisation in float64, cast to float32 (P1 / P2)."""
    x = np.asarray(values, dtype=np.float64)
    mean = x.mean(axis=1, keepdims=True)
    std = x.std(axis=1, keepdims=True)          # ddof 0, as StandardScaler
    std = np.where(std == 0.0, 1.0, std)        # sklearn _handle_zeros_in_scale
    return ((x - mean) / std).astype(np.float32)


def unpack_zscore(packed: np.ndarray, n_bits: int) -> np.ndarray:
    return zscore_rows(unpack_bits(packed, n_bits))


def u8_image_zscore(images_u8: np.ndarray) -> np.ndarray:
    """uint8 CHW depiction -> ToTensor() scaling (x/255, float32) -> per-molecule z-score."""
    x = images_u8.reshape(images_u8.shape[0], -1).astype(np.float32) / np.float32(255.0)
    return zscore_rows(x)


def standardize_chunks(x: np.ndarray, chunk: int = 100) -> np.ndarray:
    """Per-FEATURE z-score inside consecutive blocks of ``chunk`` molecules: the reference's ``standardize_features``
    (/root/reference/Descriptors/multi_input_data_preprocess_maccs_opt_IsolationForest_fixed_1.py:86-101,
    ``scaler.fit_transform(batch_features)`` per block; ``fit`` resets the scaler, so blocks are independent).  Restates
    sklearn 1.9 StandardScaler on a float32 matrix: float64 sums, the corrected two-pass variance of
    ``_incremental_mean_and_var``, ``_is_constant_feature`` -> scale 1, and the in-place float32 transform
    ``X -= float32(mean_); X /= float32(scale_)``.  Checked against sklearn itself in tests/test_oracle_pinning.py."""
    x = np.ascontiguousarray(x, dtype=np.float32)
    out = np.empty_like(x)
    eps = np.finfo(np.float64).eps
    for a in range(0, x.shape[0], chunk):
        blk = x[a:a + chunk]
        n = blk.shape[0]
        total = blk.sum(axis=0, dtype=np.float64)
        mean = total / n
        dev = blk - mean                                   # float64
        corr = dev.sum(axis=0)
        var = ((dev ** 2).sum(axis=0) - corr ** 2 / n) / n
        scale = np.sqrt(var)
        scale[var <= n * eps * var + (n * mean * eps) ** 2] = 1.0
        # sklearn 1.9 transform: X -= astype(mean_, X.dtype); X /= astype(scale_, X.dtype)  -> float32 arithmetic
        # (releases before the array-API port subtracted the float64 statistics instead: a 1-ulp difference; the reference
        # pins no sklearn version, the oracle is "sklearn as installed in this image", like torch)
        out[a:a + chunk] = (blk - mean.astype(np.float32)) / scale.astype(np.float32)
    return out


def pca_transform(x: np.ndarray, mean: np.ndarray, components: np.ndarray) -> np.ndarray:
    """(N, D) -> (N, k) in float32, ``(x - mean) @ components.T``."""
    x = np.asarray(x, dtype=np.float32)
    return (x - mean.astype(np.float32)) @ components.astype(np.float32).T
