"""The whole-model C entry points (``bbbp_fwd`` and friends, include/bbbp_b200.h) driven from Python.

This is what a host written in C / C++ / Go / Java does with the library ("A host that is not Python"):
fill a ``bbbp_model_desc``, hand over the parameters as a table of device pointers in the reference's ``state_dict()`` order
(Models/multi_input_data_regression_opt_transformer_cnn_20250113.py:69-107), let ``bbbp_model_prepare`` derive the 16-bit /
re-laid weights into a caller-owned buffer, and call ``bbbp_fwd`` with a caller-owned workspace.  The orchestration -- which
kernel, which pitch, which split-K factor -- lives in csrc/model_fwd.cu; torch is used here only to own device memory.
The scores are bit-identical to ``TransformerCnnModel.forward_groups`` (tests/test_model_gpu.py::test_c_host_forward_*).
"""
from __future__ import annotations

import ctypes

import torch

from ._lib import lib, check

PRECISION_CODES = {"bf16": 1, "fp16": 2, "strict": 3}     # BBBP_PREC_* (fp32 = 0 is not built into bbbp_fwd)
VARIANT_TCNN_20250113 = 0


class ModelDesc(ctypes.Structure):
    """``bbbp_model_desc`` (include/bbbp_b200.h)."""
    _fields_ = [("abi_version", ctypes.c_int), ("variant", ctypes.c_int), ("fingerprint_size", ctypes.c_int),
                ("precision", ctypes.c_int), ("groups", ctypes.c_int), ("seq", ctypes.c_int), ("image_is_u8", ctypes.c_int)]


def make_desc(fingerprint_size: int, precision: str, groups: int = 1, seq: int = 1, image_is_u8: bool = False) -> ModelDesc:
    if precision not in PRECISION_CODES:
        raise ValueError(f"bbbp_fwd is built for {sorted(PRECISION_CODES)}, not {precision!r}")
    return ModelDesc(lib.bbbp_abi_version(), VARIANT_TCNN_20250113, int(fingerprint_size), PRECISION_CODES[precision], int(groups),
                     int(seq), int(bool(image_is_u8)))


def param_names(desc: ModelDesc) -> list[str]:
    n = lib.bbbp_model_param_count(ctypes.byref(desc))
    if n < 0:
        check(n, "model_param_count")
    return [lib.bbbp_model_param_name(ctypes.byref(desc), i).decode() for i in range(n)]


class CHostForward:
    """``scores = CHostForward(model, precision)(fingerprint, image, groups)`` through ``bbbp_fwd``.

    ``model``: a canonical ``MixedInputModel`` on the GPU (only its ``state_dict`` tensors are used).  The prepared buffer is
    rebuilt by ``prepare()`` (call it again after the parameters change); workspaces are cached per call shape."""

    def __init__(self, model: torch.nn.Module, precision: str = "strict"):
        self.precision = precision
        state = model.state_dict()
        self.fingerprint_size = state["fingerprint_fc.0.weight"].shape[1]
        desc = make_desc(self.fingerprint_size, precision)
        names = param_names(desc)
        self._tensors = []
        for i, name in enumerate(names):
            t = state[name]
            if not t.is_cuda:
                raise RuntimeError("bbbp_b200: the model must live on the GPU (there is no CPU path)")
            if name.endswith("num_batches_tracked"):
                self._tensors.append(None)
                continue
            if t.dtype != torch.float32 or not t.is_contiguous():
                raise TypeError(f"{name}: float32 contiguous parameters expected")
            if t.numel() != lib.bbbp_model_param_numel(ctypes.byref(desc), i):
                raise ValueError(f"{name}: {t.numel()} elements, the library expects {lib.bbbp_model_param_numel(ctypes.byref(desc), i)}")
            self._tensors.append(t)
        self.device = self._tensors[0].device
        self._table = (ctypes.c_void_p * len(names))(*[None if t is None else t.data_ptr() for t in self._tensors])
        self._prepared = torch.empty((lib.bbbp_model_prepared_bytes(ctypes.byref(desc)),), device=self.device, dtype=torch.uint8)
        self._workspaces: dict = {}
        self.prepare()

    def prepare(self) -> None:
        desc = make_desc(self.fingerprint_size, self.precision)
        with torch.cuda.device(self.device):
            check(lib.bbbp_model_prepare(ctypes.byref(desc), self._table, self._prepared.data_ptr(), self._prepared.numel(),
                                         torch.cuda.current_stream().cuda_stream), "model_prepare")

    def __call__(self, fingerprint: torch.Tensor, image: torch.Tensor, groups: int = 1) -> torch.Tensor:
        rows = fingerprint.shape[0]
        if rows % groups:
            raise ValueError(f"{rows} molecules do not split into {groups} equal reference batches")
        if not (fingerprint.is_cuda and image.is_cuda):
            raise RuntimeError("bbbp_b200 models run on CUDA (sm_100a) tensors only: there is no CPU fallback")
        if fingerprint.dtype != torch.float32 or image.dtype not in (torch.float32, torch.uint8):
            raise TypeError("fingerprint float32; image float32 (standardised) or uint8 (raw depictions)")
        fingerprint, image = fingerprint.contiguous(), image.contiguous()
        if image.numel() != rows * 3 * 128 * 128:
            raise ValueError("image must hold 3 x 128 x 128 values per molecule")
        desc = make_desc(self.fingerprint_size, self.precision, groups, rows // groups, image.dtype == torch.uint8)
        key = (rows, groups, image.dtype)
        ws = self._workspaces.get(key)
        if ws is None:
            need = lib.bbbp_workspace_bytes(ctypes.byref(desc))
            if need == 0:
                check(-1, "workspace_bytes")
            if len(self._workspaces) >= 4:
                self._workspaces.pop(next(iter(self._workspaces)))
            ws = self._workspaces[key] = torch.empty((need,), device=self.device, dtype=torch.uint8)
        out = torch.empty((rows, 1), device=self.device, dtype=torch.float32)
        with torch.cuda.device(self.device):
            check(lib.bbbp_fwd(ctypes.byref(desc), fingerprint.data_ptr(), image.data_ptr(), self._table, self._prepared.data_ptr(),
                               out.data_ptr(), ws.data_ptr(), ws.numel(), torch.cuda.current_stream().cuda_stream), "fwd")
        return out
