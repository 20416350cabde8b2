"""bbbp_b200 -- B200 (sm_100a) kernels behind the reference's multi-input BBB-permeability network.

Import name: ``bbbp_b200`` (the repo-root shim ``bbbp_b200.py`` maps it onto this directory).
The package is a thin Python host over a C-ABI CUDA library (include/bbbp_b200.h); importing it
without the built library raises -- there is no CPU or eager-PyTorch fallback.
"""
from ._lib import ABI_VERSION, LIB_PATH, PROTOTYPES, last_error  # noqa: F401  (fails loudly if the .so is missing)
from . import ops  # noqa: F401
from .model import (  # noqa: F401
    AttentionFusion, BCEWithLogitsLoss, MixedInputModel, MixedInputModelBig, MixedInputModelMLP, MixedInputModelMLPMore,
    MixedInputModelMLPRdkit, MixedInputModelNoFusion, MlpModel, MSELoss, MultiHeadAttentionFusion,
    MultiModalAttentionFusion, TransformerCnnModel, VARIANTS, build, encoder_heads, zero_dropout)
from .optim import AdamW  # noqa: F401
from .train import GraphedTrainStep  # noqa: F401
from .feeder import DeviceBatchFeeder  # noqa: F401
from .data_parallel import FlatGradients  # noqa: F401
from .ensemble import OutOfFoldScores, stack_columns  # noqa: F401
from .screening import (  # noqa: F401
    SparseDepictions, average_gradients, gather_scores, pack_fingerprint_bits, partition_batches, screen, screen_library)
from .c_host import CHostForward  # noqa: F401
from .preprocess import pca_transform, standardize_chunks, unpack_zscore, u8_image_zscore  # noqa: F401

__all__ = [
    "MixedInputModel", "MixedInputModelBig", "MixedInputModelNoFusion", "MixedInputModelMLP", "MixedInputModelMLPMore",
    "MixedInputModelMLPRdkit", "MultiHeadAttentionFusion", "AttentionFusion", "MultiModalAttentionFusion", "MSELoss",
    "BCEWithLogitsLoss", "AdamW", "build", "VARIANTS", "ops", "partition_batches", "gather_scores", "screen",
    "average_gradients", "FlatGradients", "DeviceBatchFeeder", "GraphedTrainStep", "OutOfFoldScores", "stack_columns", "screen_library", "pack_fingerprint_bits", "SparseDepictions", "pca_transform", "standardize_chunks", "unpack_zscore", "u8_image_zscore", "CHostForward",
]
