"""Tensor-level wrappers over the C ABI (no autograd here; see autograd.py).

Each function takes CUDA tensors, allocates outputs/workspaces through torch's caching
allocator, and enqueues the kernel on torch's current stream.  PyTorch is plumbing only:
device memory and streams.  Nothing in this file computes with torch ops.
"""
from __future__ import annotations

import ctypes

import torch

from ._lib import lib, check

ACT_NONE, ACT_RELU, ACT_TANH = 0, 1, 2
_ACT = {None: 0, "none": 0, "relu": 1, "tanh": 2, 0: 0, 1: 1, 2: 2}

# 16-bit operand formats of the tensor-core entry points (include/bbbp_b200.h: BBBP_FMT_*)
FMT_BF16, FMT_F16 = 0, 1
_DT16 = {FMT_BF16: torch.bfloat16, FMT_F16: torch.float16}

_SM_COUNT = {}


def sm_count(device) -> int:
    idx = torch.device(device).index or 0
    if idx not in _SM_COUNT:
        _SM_COUNT[idx] = torch.cuda.get_device_properties(idx).multi_processor_count
    return _SM_COUNT[idx]


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _ptr(t):
    return None if t is None else t.data_ptr()


def _f32(t: torch.Tensor, name: str) -> torch.Tensor:
    if not t.is_cuda:
        raise RuntimeError(f"bbbp_b200: {name} must be a CUDA tensor (this package has no CPU path)")
    if t.dtype != torch.float32:
        raise TypeError(f"bbbp_b200: {name} must be float32, got {t.dtype}")
    return t if t.is_contiguous() else t.contiguous()


class _KernelTimer:
    """Optional CUDA-event brackets around named kernels (bench.py's live roofline measurement).
    Disabled by default: the wrappers then pay one set lookup."""

    def __init__(self):
        self.names, self.spans = frozenset(), {}

    def enable(self, names):
        self.names = frozenset(names)
        self.spans = {n: [] for n in names}

    def disable(self):
        self.names = frozenset()

    def start(self, name):
        if name not in self.names:
            return None
        e = torch.cuda.Event(enable_timing=True)
        e.record()
        return e

    def stop(self, name, e0, units):
        if e0 is not None:
            e1 = torch.cuda.Event(enable_timing=True)
            e1.record()
            self.spans[name].append((e0, e1, units))

    def collect(self, name):
        torch.cuda.synchronize()
        spans = self.spans.get(name, [])
        return [a.elapsed_time(b) for a, b, _ in spans], [u for _, _, u in spans]


KERNEL_TIMER = _KernelTimer()


def require_device() -> None:
    check(lib.bbbp_device_check(), "device_check")


# ---- dense ---------------------------------------------------------------------------------------------------------
def gemm_f32(a, b, trans_a=False, trans_b=False, bias=None, act=None, out=None, accumulate=False, split_k=1):
    """act(op(a) @ op(b) + bias) with row-major 2-D float32 operands (pitches = tensor strides)."""
    assert a.dim() == 2 and b.dim() == 2 and a.stride(1) == 1 and b.stride(1) == 1
    M, K = (a.shape[1], a.shape[0]) if trans_a else a.shape
    Kb, N = (b.shape[1], b.shape[0]) if trans_b else b.shape
    assert K == Kb, (a.shape, b.shape, trans_a, trans_b)
    if out is None:
        out = torch.empty((M, N), device=a.device, dtype=torch.float32)
    ws = None
    if split_k > 1:
        ws = torch.empty((split_k * M * N,), device=a.device, dtype=torch.float32)
    elif split_k == 0:          # latency mode (training): kernel and K partition chosen by the library
        need = lib.bbbp_gemm_f32_auto_workspace(M, N, K)
        if need:
            ws = torch.empty((need // 4,), device=a.device, dtype=torch.float32)
    check(lib.bbbp_gemm_f32(int(trans_a), int(trans_b), M, N, K, a.data_ptr(), a.stride(0), b.data_ptr(), b.stride(0),
                            out.data_ptr(), out.stride(0), _ptr(bias), _ACT[act], int(accumulate), split_k, _ptr(ws),
                            0 if ws is None else ws.numel() * 4, _stream()), "gemm_f32")
    return out


def fixed_split_k(K: int) -> int:
    """Split-K factor as a function of K ONLY.  Each output row then sees the same reduction order whatever the
    number of rows in the call, so a molecule's score does not depend on how many reference batches share a launch
    or on how batches are sharded over GPUs (bit-identical 1/2/4/8-GPU screening)."""
    return 1 if K < 8192 else min(8, K // 2048)


def fixed_split_k_strict(K: int) -> int:
    """Strict mode: the tensor core's fp32 accumulator is not a round-to-nearest one, and its error grows with the length
    of an accumulation chain (measured on Linear(65536, 128), tests/strict_error_budget.py: relative error 2.6e-4 with one
    chain of 1 024 K-blocks, 3.7e-5 with 8 chains, 1.2e-5 with 32, 8e-6 with 128).  More split-K partials -- summed in fp32
    by the finish kernel -- shorten every chain: 32 K-blocks of 64 at most.  Still a function of K only."""
    return 1 if K < 4096 else min(64, K // 2048)


def fixed_split_k_f32(K: int) -> int:
    """Same contract for the CUDA-core fp32 GEMM (64-wide tiles, used at small batch where tiles are few)."""
    return 1 if K < 1024 else min(64, K // 256)


def tile_split_k(M: int, N: int, K: int, device) -> int:
    """Backward GEMMs carry no bit-identity contract: split K until the grid fills the GPU."""
    tiles = -(-M // 64) * -(-N // 64)
    return max(1, min(-(-2 * sm_count(device) // tiles), K // 128))


def cast_bf16(x: torch.Tensor, ld: int | None = None, out: torch.Tensor | None = None) -> torch.Tensor:
    """(rows, cols) float32 -> (rows, ld) bfloat16 with zero fill of the pad columns; ld multiple of 8.
    ``out`` may be a pitched 2-D bf16 view (its row stride is the destination pitch)."""
    rows, cols = x.shape
    if out is None:
        if ld is None:
            ld = -(-cols // 8) * 8
        out = torch.empty((rows, ld), device=x.device, dtype=torch.bfloat16)
        pad_to, pitch = ld, ld
    else:
        pad_to, pitch = out.shape[1], out.stride(0)
    check(lib.bbbp_cast_bf16(x.data_ptr(), x.stride(0), out.data_ptr(), pitch, rows, cols, pad_to, _stream()), "cast_bf16")
    return out


def cast16(x: torch.Tensor, fmt: int = FMT_BF16, ld: int | None = None, want_lo: bool = False, out: torch.Tensor | None = None,
           sub: torch.Tensor | None = None):
    """(rows, cols) float32 -> (rows, ld) 16-bit rows in format ``fmt`` (pad columns zero, ld a multiple of 8).  With
    ``want_lo`` also the low part ``rn(x - hi)``: returns (hi, lo | None).  128-bit stores (one thread = 8 columns).
    ``out``: a pitched 2-D destination view for hi (row stride = pitch, width = pad-to, both multiples of 8).
    ``sub``: a (cols,) float32 row subtracted from every row before the split (fused centring)."""
    rows, cols = x.shape
    if out is not None:
        assert not want_lo and out.shape[0] == rows
        check(lib.bbbp_cast16(fmt, x.data_ptr(), x.stride(0), _ptr(sub), out.data_ptr(), None, out.stride(0), rows, cols,
                              out.shape[1], _stream()), "cast16")
        return out, None
    ld = -(-cols // 8) * 8 if ld is None else ld
    hi = torch.empty((rows, ld), device=x.device, dtype=_DT16[fmt])
    lo = torch.empty((rows, ld), device=x.device, dtype=_DT16[fmt]) if want_lo else None
    check(lib.bbbp_cast16(fmt, x.data_ptr(), x.stride(0), _ptr(sub), hi.data_ptr(), _ptr(lo), ld, rows, cols, ld, _stream()),
          "cast16")
    return hi, lo


def standardize_chunks(x: torch.Tensor, chunk_rows: int = 100, out: torch.Tensor | None = None) -> torch.Tensor:
    """Per-feature z-score inside consecutive blocks of ``chunk_rows`` rows (StandardScaler().fit_transform per block)."""
    assert x.dim() == 2 and x.dtype == torch.float32 and x.stride(1) == 1
    out = torch.empty_like(x) if out is None else out
    check(lib.bbbp_standardize_chunks_f32(x.data_ptr(), x.stride(0), out.data_ptr(), out.stride(0), x.shape[0], x.shape[1],
                                          int(chunk_rows), _stream()), "standardize_chunks")
    return out


def gemm_bf16(a16, K, w16, N, bias=None, residual=None, act=None, out_f32=True, out_bf16=False, split_k=1,
              ld_out=None, ld_out16=None, fmt=FMT_BF16, a_lo=None, w_lo=None, out16_lo=None, pre_add=None):
    """act(a16[:, :K] @ w16[:N, :K]^T + bias) (+ residual) on the tcgen05 path.  Returns (f32 | None, 16-bit | None);
    outputs are (M, ld) buffers whose first N columns are the result (16-bit pad columns are zero).  ``fmt``: operand
    format; ``a_lo`` / ``w_lo``: optional low parts (one more MMA each per K step); ``out16_lo`` (True / False, default
    None): return a 3-tuple (f32, hi, lo) whose lo is the low part of the 16-bit output when True, None when False.
    ``pre_add``: an (M, >= N) float32 addend applied before the activation (per-row bias; excludes ``residual``)."""
    M = a16.shape[0]
    ld_out = N if ld_out is None else ld_out
    ld16 = -(-N // 8) * 8 if ld_out16 is None else ld_out16
    o32 = torch.empty((M, ld_out), device=a16.device, dtype=torch.float32) if out_f32 else None
    o16 = torch.empty((M, ld16), device=a16.device, dtype=_DT16[fmt]) if out_bf16 else None
    o16lo = torch.empty((M, ld16), device=a16.device, dtype=_DT16[fmt]) if (out_bf16 and out16_lo) else None
    ws_bytes = lib.bbbp_gemm_bf16_workspace(M, N, split_k)
    ws = torch.empty((ws_bytes,), device=a16.device, dtype=torch.uint8) if ws_bytes else None
    if o16 is not None and ws_bytes and ld16 > N:
        fill_zero(o16[:, N:])          # the split-K finish kernel writes only the N result columns
        if o16lo is not None:
            fill_zero(o16lo[:, N:])
    if pre_add is not None:
        assert residual is None and pre_add.dtype == torch.float32 and pre_add.shape[0] == M and pre_add.stride(1) == 1
        check(lib.bbbp_gemm16_pre(fmt, M, N, K, a16.data_ptr(), _ptr(a_lo), a16.stride(0), w16.data_ptr(), _ptr(w_lo),
                                  w16.stride(0), _ptr(bias), pre_add.data_ptr(), pre_add.stride(0), _ptr(o32), ld_out, _ptr(o16),
                                  _ptr(o16lo), ld16, _ACT[act], split_k, _ptr(ws), ws_bytes, _stream()), "gemm16_pre")
        return (o32, o16, o16lo) if out16_lo is not None else (o32, o16)
    check(lib.bbbp_gemm16(fmt, M, N, K, a16.data_ptr(), _ptr(a_lo), a16.stride(0), w16.data_ptr(), _ptr(w_lo), w16.stride(0),
                          _ptr(bias), _ptr(residual), 0 if residual is None else residual.stride(0), _ptr(o32), ld_out,
                          _ptr(o16), _ptr(o16lo), ld16, _ACT[act], split_k, _ptr(ws), ws_bytes, _stream()), "gemm16")
    return (o32, o16, o16lo) if out16_lo is not None else (o32, o16)


def gemm16_tn(a16, w16, M, N, K, trans_a=False, trans_w=False, bias=None, act=None, split_k=1, fmt=FMT_BF16, out16=False,
              out_f32=True):
    """act(opA(a16) @ opW(w16)^T + bias) with operands read in place: trans_a -> a16 is stored (K, M); trans_w -> w16 is
    stored (K, N) (row pitches = tensor strides).  Returns (f32 | None, 16-bit | None)."""
    o32 = torch.empty((M, N), device=a16.device, dtype=torch.float32) if out_f32 else None
    ld16 = -(-N // 8) * 8
    o16 = torch.empty((M, ld16), device=a16.device, dtype=_DT16[fmt]) if out16 else None
    ws_bytes = lib.bbbp_gemm_bf16_workspace(M, N, split_k)
    ws = torch.empty((ws_bytes,), device=a16.device, dtype=torch.uint8) if ws_bytes else None
    if o16 is not None and ws_bytes and ld16 > N:
        fill_zero(o16[:, N:])
    check(lib.bbbp_gemm16_tn(fmt, int(trans_a), int(trans_w), M, N, K, a16.data_ptr(), a16.stride(0), w16.data_ptr(), w16.stride(0),
                             _ptr(bias), _ptr(o32), N, _ptr(o16), ld16, _ACT[act], split_k, _ptr(ws), ws_bytes, _stream()),
          "gemm16_tn")
    return o32, o16


def fill_zero(t: torch.Tensor) -> torch.Tensor:
    """Zero a (possibly pitched) 2-D view or a contiguous tensor with the library's fill kernel."""
    if t.numel() == 0:
        return t
    if t.dim() == 2 and not t.is_contiguous():
        rows, cols, pitch = t.shape[0], t.shape[1], t.stride(0)
    else:
        assert t.is_contiguous()
        rows, cols, pitch = 1, t.numel(), t.numel()
    esz = t.element_size()
    check(lib.bbbp_fill_zero(t.data_ptr(), rows, cols * esz, pitch * esz, _stream()), "fill_zero")
    return t


def gemm_bf16_batched(batches, M, N, K, a16, lda, a_bs, w16, ldw, w_bs, out_bf16=True, out_f32=False, ld_out16=None,
                      ld_out=None, fmt=FMT_BF16):
    """out[b] = A[b] @ W[b]^T for b < batches; outputs are (batches*M, ld) with batch b at rows [b*M, (b+1)*M)."""
    ld16 = -(-N // 8) * 8 if ld_out16 is None else ld_out16
    ld32 = N if ld_out is None else ld_out
    o32 = torch.empty((batches * M, ld32), device=a16.device, dtype=torch.float32) if out_f32 else None
    o16 = torch.empty((batches * M, ld16), device=a16.device, dtype=_DT16[fmt]) if out_bf16 else None
    check(lib.bbbp_gemm16_batched(fmt, batches, M, N, K, a16.data_ptr(), lda, a_bs, w16.data_ptr(), ldw, w_bs, _ptr(o32), ld32,
                                  M * ld32, _ptr(o16), ld16, M * ld16, _stream()), "gemm16_batched")
    return o32, o16


def softmax_rows_scaled_bf16(scores: torch.Tensor, cols: int, scale: float, fmt=FMT_BF16) -> torch.Tensor:
    """(rows, ld) fp32 logits -> (rows, ceil8(cols)) 16-bit softmax(scale * logits) with zero pad columns."""
    rows = scores.shape[0]
    ldp = -(-cols // 8) * 8
    p = torch.empty((rows, ldp), device=scores.device, dtype=_DT16[fmt])
    check(lib.bbbp_softmax_rows_scaled16(fmt, scores.data_ptr(), scores.stride(0), p.data_ptr(), ldp, rows, cols, float(scale),
                                         _stream()), "softmax_rows_scaled16")
    return p


def attention_scores_softmax_bf16(q16, k16, ld, groups, seq, head_dim, scale, fmt=FMT_BF16):
    """softmax(scale * Q K^T) per group as bf16 (groups*seq, ldp); q16 / k16 are views into the packed qkv buffer.
    seq <= 256: one GEMM with the softmax in its TMEM epilogue.  Wider scopes: batched GEMM to fp32 logits, then a
    row-softmax kernel (the S x S logits are materialised: 4*S*S bytes per group)."""
    if seq > 256:
        ld_s = -(-seq // 4) * 4
        s32, _ = gemm_bf16_batched(groups, seq, seq, head_dim, q16, ld, seq * ld, k16, ld, seq * ld, out_bf16=False,
                                   out_f32=True, ld_out=ld_s, fmt=fmt)
        return softmax_rows_scaled_bf16(s32, seq, scale, fmt)
    ldp = -(-seq // 8) * 8
    p = torch.empty((groups * seq, ldp), device=q16.device, dtype=_DT16[fmt])
    t0 = KERNEL_TIMER.start("attn_scores")
    check(lib.bbbp_attention_scores_softmax16(fmt, groups, seq, head_dim, q16.data_ptr(), ld, k16.data_ptr(), ld, seq * ld,
                                              float(scale), p.data_ptr(), ldp, _stream()), "attention_scores_softmax")
    KERNEL_TIMER.stop("attn_scores", t0, groups * seq)
    return p


def attention_flash16(q16, k16, ld, groups, seq, head_dim, scale, vt, ld_vt, fmt=FMT_BF16, ld_out=None):
    """softmax(scale * Q K^T) V per group with the streaming tcgen05 kernel (any seq; one head, head_dim <= 192).  q16 / k16:
    views into the packed qkv buffer (row pitch ``ld``); vt: (groups, head_dim, ld_vt) transposed values.  Returns
    (groups*seq, ld_out) 16-bit rows, pad columns zero."""
    ld_out = -(-head_dim // 8) * 8 if ld_out is None else ld_out
    out = torch.empty((groups * seq, ld_out), device=q16.device, dtype=_DT16[fmt])
    if ld_out > -(-head_dim // 16) * 16:
        fill_zero(out[:, -(-head_dim // 16) * 16:])
    t0 = KERNEL_TIMER.start("attn_flash")
    check(lib.bbbp_attention_flash16(fmt, groups, seq, head_dim, q16.data_ptr(), ld, k16.data_ptr(), ld, seq * ld, vt.data_ptr(),
                                     ld_vt, head_dim * ld_vt, float(scale), out.data_ptr(), ld_out, seq * ld_out, _stream()),
          "attention_flash16")
    KERNEL_TIMER.stop("attn_flash", t0, groups * seq)
    return out


def attention_flash_proj_ln16(q16, k16, ld, groups, seq, head_dim, scale, vt, ld_vt, w_out16, b_out, residual, gamma, beta,
                              eps=1e-5, ld_y=None, ld16=0, fmt=FMT_BF16):
    """LayerNorm(residual + softmax(scale Q K^T) V W_out^T + b_out): the streaming attention kernel with out_proj + residual +
    norm1 fused into its tail.  Returns ((groups*seq, ld_y) fp32, (groups*seq, ld16) 16-bit | None)."""
    rows = groups * seq
    ld_y = head_dim if ld_y is None else ld_y
    y = torch.empty((rows, ld_y), device=q16.device, dtype=torch.float32)
    y16 = torch.empty((rows, ld16), device=q16.device, dtype=_DT16[fmt]) if ld16 else None
    t0 = KERNEL_TIMER.start("attn_flash")
    check(lib.bbbp_attention_flash_proj_ln16(fmt, groups, seq, head_dim, q16.data_ptr(), ld, k16.data_ptr(), ld, seq * ld, vt.data_ptr(),
                                             ld_vt, head_dim * ld_vt, float(scale), w_out16.data_ptr(), w_out16.stride(0),
                                             b_out.data_ptr(), residual.data_ptr(), residual.stride(0), gamma.data_ptr(),
                                             beta.data_ptr(), float(eps), y.data_ptr(), ld_y, _ptr(y16), ld16, _stream()),
          "attention_flash_proj_ln16")
    KERNEL_TIMER.stop("attn_flash", t0, rows)
    return y, y16


def transpose_bf16(src, batches, rows, cols, ld_src, src_bs, ld_dst):
    """(batches, rows, cols) pitched bf16 -> (batches, cols, ld_dst) with zero fill of columns >= rows."""
    dst = torch.empty((batches, cols, ld_dst), device=src.device, dtype=src.dtype)
    check(lib.bbbp_transpose_bf16(batches, rows, cols, src.data_ptr(), ld_src, src_bs, dst.data_ptr(), ld_dst, cols * ld_dst,
                                  _stream()), "transpose_bf16")
    return dst


# ---- image branch -----------------------------------------------------------------------------------------------------
def conv3x3(x, w, b, pool=True, want_argmax=False):
    N, Cin, H, W = x.shape
    Cout = w.shape[0]
    if pool:
        y = torch.empty((N, Cout, H // 2, W // 2), device=x.device, dtype=torch.float32)
        arg = torch.empty(y.shape, device=x.device, dtype=torch.uint8) if want_argmax else None
    else:
        y = torch.empty((N, Cout, H, W), device=x.device, dtype=torch.float32)
        arg = None
    name = "conv1" if Cin == 3 else "conv2" if (Cin, Cout) == (32, 64) else "conv"
    t0 = KERNEL_TIMER.start(name)
    check(lib.bbbp_conv3x3_f32(x.data_ptr(), w.data_ptr(), _ptr(b), y.data_ptr(), _ptr(arg), N, Cin, Cout, H, W,
                               int(pool), _stream()), "conv3x3_f32")
    KERNEL_TIMER.stop(name, t0, N)
    return y, arg


def relu_pool_bwd(dy, y, argmax, H, W):
    N, C = y.shape[:2]
    dpre = torch.empty((N, C, H, W), device=y.device, dtype=torch.float32)
    check(lib.bbbp_relu_pool_bwd_f32(dy.data_ptr(), y.data_ptr(), argmax.data_ptr(), dpre.data_ptr(), N, C, H, W,
                                     _stream()), "relu_pool_bwd")
    return dpre


def conv3x3_wgrad(dpre, x, Cout, Cin):
    N, _, H, W = x.shape
    dw = torch.empty((Cout, Cin, 3, 3), device=x.device, dtype=torch.float32)
    db = torch.empty((Cout,), device=x.device, dtype=torch.float32)
    ws = torch.empty((lib.bbbp_conv3x3_wgrad_workspace(N, Cin, Cout, H, W) // 4,), device=x.device, dtype=torch.float32)
    check(lib.bbbp_conv3x3_wgrad_f32(dpre.data_ptr(), x.data_ptr(), dw.data_ptr(), db.data_ptr(), N, Cin, Cout, H, W,
                                     ws.data_ptr(), ws.numel() * 4, _stream()), "conv3x3_wgrad")
    return dw, db


def conv3x3_flip_weights(w):
    Cout, Cin = w.shape[:2]
    wt = torch.empty((Cin, Cout, 3, 3), device=w.device, dtype=torch.float32)
    check(lib.bbbp_conv3x3_flip_weights_f32(w.data_ptr(), wt.data_ptr(), Cin, Cout, _stream()), "flip_weights")
    return wt


# ---- tcgen05 image branch (inference) ---------------------------------------------------------------------------------
def conv3x3_prepare_bf16(w: torch.Tensor, fmt=FMT_BF16) -> torch.Tensor:
    Cout, Cin = w.shape[:2]
    n = lib.bbbp_conv3x3_prepared_bytes(Cin, Cout)
    wp = torch.empty((n,), device=w.device, dtype=torch.uint8)
    check(lib.bbbp_conv3x3_prepare16(fmt, w.data_ptr(), wp.data_ptr(), Cin, Cout, _stream()), "conv3x3_prepare16")
    return wp


def image_to_nhwc8_bf16(img: torch.Tensor, C=3, H=128, W=128) -> torch.Tensor:
    N = img.numel() // (C * H * W)
    out = torch.empty((N, H, W, 8), device=img.device, dtype=torch.bfloat16)
    check(lib.bbbp_image_to_nhwc8_bf16(img.data_ptr(), out.data_ptr(), N, C, H, W, _stream()), "image_to_nhwc8")
    return out


def conv3x3_relu_pool_bf16(x_nhwc: torch.Tensor, wprep: torch.Tensor, bias: torch.Tensor, Cout: int, fmt=FMT_BF16,
                           x_lo: torch.Tensor | None = None):
    """tcgen05 conv3x3 + bias + ReLU + 2x2 max-pool on NHWC 16-bit activations.  ``x_lo`` (strict mode): the low part of a
    (hi, lo) input pair; the output is then a (hi, lo) pair as well -> returns (y, y_lo)."""
    N, H, W, Cin_pad = x_nhwc.shape
    y = torch.empty((N, H // 2, W // 2, Cout), device=x_nhwc.device, dtype=_DT16[fmt])
    y_lo = torch.empty_like(y) if x_lo is not None else None
    name = "conv1" if Cin_pad == 8 else "conv2"
    t0 = KERNEL_TIMER.start(name)
    check(lib.bbbp_conv3x3_relu_pool16(fmt, 2 if x_lo is not None else 1, x_nhwc.data_ptr(), _ptr(x_lo), wprep.data_ptr(),
                                       bias.data_ptr(), y.data_ptr(), _ptr(y_lo), N, Cin_pad, Cout, H, W, _stream()),
          "conv3x3_relu_pool16")
    KERNEL_TIMER.stop(name, t0, N)
    return y if x_lo is None else (y, y_lo)


def decode_sparse_depictions(mask: torch.Tensor, values: torch.Tensor, offsets: torch.Tensor, out: torch.Tensor | None = None):
    """(n, 2048) uint8 bit masks + RGB triples of the non-white pixels + (n + 1,) int64 running counts -> (n, 3, 128, 128) uint8."""
    n = mask.shape[0]
    assert mask.dtype == torch.uint8 and mask.shape[1] == 2048 and mask.is_contiguous() and offsets.dtype == torch.int64
    if out is None:
        out = torch.empty((n, 3, 128, 128), device=mask.device, dtype=torch.uint8)
    check(lib.bbbp_decode_sparse_depictions_u8(mask.data_ptr(), values.data_ptr(), offsets.data_ptr(), out.data_ptr(), n, _stream()),
          "decode_sparse_depictions")
    return out


def u8_image_stats(img_u8: torch.Tensor) -> torch.Tensor:
    rows = img_u8.shape[0]
    flat = img_u8.reshape(rows, -1)
    assert flat.dtype == torch.uint8 and flat.is_contiguous()
    stats = torch.empty((rows, 2), device=img_u8.device, dtype=torch.float32)
    check(lib.bbbp_u8_image_stats_f32(flat.data_ptr(), stats.data_ptr(), rows, flat.shape[1], _stream()), "u8_image_stats")
    return stats


def conv1_from_image_bf16(img: torch.Tensor, wprep: torch.Tensor, bias: torch.Tensor, stats=None, H=128, W=128,
                          fmt=FMT_BF16, split=False):
    """conv1 + ReLU + pool straight from the planar (N, 3, H, W) input: fp32 (standardised) or uint8 (+ stats).
    ``split`` (strict mode): the producers stage the image as a (hi, lo) pair and the output is a (hi, lo) pair."""
    N = img.numel() // (3 * H * W)
    y = torch.empty((N, H // 2, W // 2, 32), device=img.device, dtype=_DT16[fmt])
    y_lo = torch.empty_like(y) if split else None
    is_u8 = img.dtype == torch.uint8
    assert is_u8 or img.dtype == torch.float32
    t0 = KERNEL_TIMER.start("conv1")
    check(lib.bbbp_conv1_from_image16(fmt, 2 if split else 1, img.data_ptr(), int(is_u8), _ptr(stats), wprep.data_ptr(),
                                      bias.data_ptr(), y.data_ptr(), _ptr(y_lo), N, H, W, _stream()), "conv1_from_image16")
    KERNEL_TIMER.stop("conv1", t0, N)
    return (y, y_lo) if split else y


def conv1_from_image_c64(img: torch.Tensor, wprep: torch.Tensor, bias: torch.Tensor, H=128, W=128) -> torch.Tensor:
    """Conv2d(3, 64) + ReLU + MaxPool2d(2) of the big variant straight from the planar fp32 (N, 3, H, W) input -> bf16 NHWC."""
    N = img.numel() // (3 * H * W)
    assert img.dtype == torch.float32
    y = torch.empty((N, H // 2, W // 2, 64), device=img.device, dtype=torch.bfloat16)
    t0 = KERNEL_TIMER.start("conv1")
    check(lib.bbbp_conv1_from_image_c64_bf16(img.data_ptr(), wprep.data_ptr(), bias.data_ptr(), y.data_ptr(), N, H, W, _stream()),
          "conv1_from_image_c64")
    KERNEL_TIMER.stop("conv1", t0, N)
    return y


# ---- background-referenced strict mode of the image branch (conv_umma.cu, BG = 1) ---------------------------------------
def image_background(img: torch.Tensor, stats=None, H=128, W=128) -> torch.Tensor:
    """(N, 4) float32: the background value of every image per channel (column 3 is zero); uint8 images are normalised
    with ``stats`` exactly as the first layer's producers do."""
    N = img.numel() // (3 * H * W)
    is_u8 = img.dtype == torch.uint8
    assert is_u8 or img.dtype == torch.float32
    bg = torch.empty((N, 4), device=img.device, dtype=torch.float32)
    check(lib.bbbp_image_background(img.data_ptr(), int(is_u8), _ptr(stats), bg.data_ptr(), N, H, W, _stream()), "image_background")
    return bg


def fc_weight_channel_sums(w: torch.Tensor, C: int, HW: int) -> torch.Tensor:
    """(rows, C) float32: sum over the HW positions of w[o, c*HW:(c+1)*HW] (nn.Flatten's (C, H, W) order; HW = 9: a 3x3
    convolution weight summed over its taps)."""
    rows = w.shape[0]
    w = _f32(w, "w")
    out = torch.empty((rows, C), device=w.device, dtype=torch.float32)
    check(lib.bbbp_fc_weight_channel_sums(w.data_ptr(), out.data_ptr(), rows, C, HW, _stream()), "fc_weight_channel_sums")
    return out


def bg_layer(wsum: torch.Tensor, bias, bg_in: torch.Tensor, fmt: int = -1, want_neg16: bool = False, table: bool = True):
    """One step of the background chain (bbbp_bg_layer).  ``table``: returns (tab (N, 2, Cout) = {T, bg_out}, neg16 | None)
    for a convolution layer; otherwise just T (N, Cout) (the Linear's per-row addend)."""
    Cout, Cin = wsum.shape
    N = bg_in.shape[0]
    if not table:
        out = torch.empty((N, Cout), device=wsum.device, dtype=torch.float32)
        check(lib.bbbp_bg_layer(wsum.data_ptr(), _ptr(bias), bg_in.data_ptr(), bg_in.stride(0), Cin, Cout, out.data_ptr(), Cout,
                                None, 0, None, -1, N, _stream()), "bg_layer")
        return out
    tab = torch.empty((N, 2, Cout), device=wsum.device, dtype=torch.float32)
    neg = torch.empty((N, Cout), device=wsum.device, dtype=_DT16[fmt]) if want_neg16 else None
    check(lib.bbbp_bg_layer(wsum.data_ptr(), _ptr(bias), bg_in.data_ptr(), bg_in.stride(0), Cin, Cout, tab.data_ptr(), 2 * Cout,
                            tab[:, 1].data_ptr(), 2 * Cout, _ptr(neg), fmt, N, _stream()), "bg_layer")
    return tab, neg


def conv1_from_image_bg(img: torch.Tensor, wprep: torch.Tensor, stats, bg_in, tab, H=128, W=128, fmt=FMT_F16, split=True):
    """First layer on x - bg_in (staged as a (hi, lo) pair when ``split``); returns maxpool(relu(conv1(x))) - tab[:, 1] as
    ONE NHWC 16-bit tensor."""
    N = img.numel() // (3 * H * W)
    y = torch.empty((N, H // 2, W // 2, 32), device=img.device, dtype=_DT16[fmt])
    is_u8 = img.dtype == torch.uint8
    t0 = KERNEL_TIMER.start("conv1")
    check(lib.bbbp_conv1_from_image_bg16(fmt, 2 if split else 1, img.data_ptr(), int(is_u8), _ptr(stats), wprep.data_ptr(),
                                         bg_in.data_ptr(), tab.data_ptr(), y.data_ptr(), N, H, W, _stream()),
          "conv1_from_image_bg16")
    KERNEL_TIMER.stop("conv1", t0, N)
    return y


def conv3x3_relu_pool_bg(x_nhwc: torch.Tensor, wprep: torch.Tensor, neg_bg_in, tab, Cout: int, fmt=FMT_F16):
    """Second layer on background-referenced NHWC input (``neg_bg_in``: what its padding holds); returns
    maxpool(relu(conv(x))) - tab[:, 1]."""
    N, H, W, Cin_pad = x_nhwc.shape
    y = torch.empty((N, H // 2, W // 2, Cout), device=x_nhwc.device, dtype=_DT16[fmt])
    t0 = KERNEL_TIMER.start("conv2")
    check(lib.bbbp_conv3x3_relu_pool_bg16(fmt, x_nhwc.data_ptr(), wprep.data_ptr(), neg_bg_in.data_ptr(), tab.data_ptr(), y.data_ptr(),
                                          N, Cin_pad, Cout, H, W, _stream()), "conv3x3_relu_pool_bg16")
    KERNEL_TIMER.stop("conv2", t0, N)
    return y


def fc_weight_to_hwc_bf16(w: torch.Tensor, C: int, HW: int, fmt=FMT_BF16) -> torch.Tensor:
    rows = w.shape[0]
    out = torch.empty((rows, C * HW), device=w.device, dtype=_DT16[fmt])
    check(lib.bbbp_fc_weight_to_hwc16(fmt, w.data_ptr(), out.data_ptr(), rows, C, HW, _stream()), "fc_weight_to_hwc")
    return out


def im2col3x3_bf16(x_nhwc: torch.Tensor) -> torch.Tensor:
    """(N, H, W, C) bf16 -> (N*H*W, 9*C) bf16 rows in (tap, channel) order, zero padding at the image border."""
    N, H, W, C = x_nhwc.shape
    out = torch.empty((N * H * W, 9 * C), device=x_nhwc.device, dtype=torch.bfloat16)
    check(lib.bbbp_im2col3x3_bf16(x_nhwc.data_ptr(), out.data_ptr(), N, H, W, C, _stream()), "im2col3x3_bf16")
    return out


def conv3x3_gemm16(x_nhwc: torch.Tensor, w_taps: torch.Tensor, bias, act="relu", fmt=FMT_BF16) -> torch.Tensor:
    """act(conv3x3(x) + bias) on NHWC 16-bit activations as an implicit GEMM (no im2col matrix); C % 64 == 0.
    w_taps: (Cout, 9*C) in (tap, channel) order (conv3x3_weight_im2col_bf16)."""
    N, H, W, C = x_nhwc.shape
    Cout = w_taps.shape[0]
    y = torch.empty((N, H, W, Cout), device=x_nhwc.device, dtype=x_nhwc.dtype)
    t0 = KERNEL_TIMER.start("conv_gemm")
    check(lib.bbbp_conv3x3_gemm16(fmt, x_nhwc.data_ptr(), N, H, W, C, w_taps.data_ptr(), Cout, _ptr(bias), _ACT[act], y.data_ptr(),
                                  _stream()), "conv3x3_gemm16")
    KERNEL_TIMER.stop("conv_gemm", t0, N)
    return y


def maxpool2x2_nhwc_bf16(x_nhwc: torch.Tensor) -> torch.Tensor:
    N, H, W, C = x_nhwc.shape
    y = torch.empty((N, H // 2, W // 2, C), device=x_nhwc.device, dtype=torch.bfloat16)
    check(lib.bbbp_maxpool2x2_nhwc_bf16(x_nhwc.data_ptr(), y.data_ptr(), N, H, W, C, _stream()), "maxpool2x2_nhwc_bf16")
    return y


def conv3x3_weight_im2col_bf16(w: torch.Tensor, c_pad: int | None = None) -> torch.Tensor:
    Cout, Cin = w.shape[:2]
    c_pad = -(-Cin // 8) * 8 if c_pad is None else c_pad
    out = torch.empty((Cout, 9 * c_pad), device=w.device, dtype=torch.bfloat16)
    check(lib.bbbp_conv3x3_weight_im2col_bf16(w.data_ptr(), out.data_ptr(), Cin, c_pad, Cout, _stream()), "conv3x3_weight_im2col")
    return out


# ---- mixed-precision training path of the image branch (conv_train.cu) ---------------------------------------------------
def image_to_nhwc8_16(img: torch.Tensor, fmt=FMT_BF16, C=3, H=128, W=128) -> torch.Tensor:
    N = img.numel() // (C * H * W)
    out = torch.empty((N, H, W, 8), device=img.device, dtype=_DT16[fmt])
    check(lib.bbbp_image_to_nhwc8_16(fmt, img.data_ptr(), out.data_ptr(), N, C, H, W, _stream()), "image_to_nhwc8_16")
    return out


def im2col3x3_16(x_nhwc: torch.Tensor) -> torch.Tensor:
    """(N, H, W, C) 16-bit (either format) -> (N*H*W, 9*C) rows in (tap, channel) order, zero padding at the border."""
    N, H, W, C = x_nhwc.shape
    out = torch.empty((N * H * W, 9 * C), device=x_nhwc.device, dtype=x_nhwc.dtype)
    check(lib.bbbp_im2col3x3_bf16(x_nhwc.data_ptr(), out.data_ptr(), N, H, W, C, _stream()), "im2col3x3")
    return out


def maxpool2x2_argmax_nhwc16(x_nhwc: torch.Tensor, fmt=FMT_BF16):
    N, H, W, C = x_nhwc.shape
    y = torch.empty((N, H // 2, W // 2, C), device=x_nhwc.device, dtype=x_nhwc.dtype)
    arg = torch.empty((N, H // 2, W // 2, C), device=x_nhwc.device, dtype=torch.uint8)
    check(lib.bbbp_maxpool2x2_argmax_nhwc16(fmt, x_nhwc.data_ptr(), y.data_ptr(), arg.data_ptr(), N, H, W, C, _stream()),
          "maxpool2x2_argmax_nhwc16")
    return y, arg


def unpool_relu_nhwc16(dy: torch.Tensor, y: torch.Tensor, arg: torch.Tensor, fmt=FMT_BF16, want_masked=True):
    """dy: fp32 (N, H/2, W/2, C) contiguous; returns (dpre 16-bit (N, H, W, C), dy * (y > 0) fp32 | None)."""
    N, OH, OW, C = y.shape
    dpre = torch.empty((N, 2 * OH, 2 * OW, C), device=y.device, dtype=y.dtype)
    dym = torch.empty((N, OH, OW, C), device=y.device, dtype=torch.float32) if want_masked else None
    check(lib.bbbp_unpool_relu_nhwc16(fmt, dy.data_ptr(), y.data_ptr(), arg.data_ptr(), dpre.data_ptr(), _ptr(dym), N, 2 * OH, 2 * OW,
                                      C, _stream()), "unpool_relu_nhwc16")
    return dpre, dym


def conv3x3_weight_im2col16(w: torch.Tensor, c_pad: int, fmt=FMT_BF16) -> torch.Tensor:
    Cout, Cin = w.shape[:2]
    out = torch.empty((Cout, 9 * c_pad), device=w.device, dtype=_DT16[fmt])
    check(lib.bbbp_conv3x3_weight_im2col16(fmt, w.data_ptr(), out.data_ptr(), Cin, c_pad, Cout, _stream()), "conv3x3_weight_im2col16")
    return out


def conv3x3_wgrad_from_im2col(g: torch.Tensor, Cin: int, c_pad: int) -> torch.Tensor:
    Cout = g.shape[0]
    dw = torch.empty((Cout, Cin, 3, 3), device=g.device, dtype=torch.float32)
    check(lib.bbbp_conv3x3_wgrad_from_im2col_f32(g.data_ptr(), dw.data_ptr(), Cin, c_pad, Cout, _stream()), "wgrad_from_im2col")
    return dw


def conv3x3_weight_dgrad16(w: torch.Tensor, cin_pad: int, fmt=FMT_BF16) -> torch.Tensor:
    Cout, Cin = w.shape[:2]
    out = torch.empty((cin_pad, 9 * Cout), device=w.device, dtype=_DT16[fmt])
    check(lib.bbbp_conv3x3_weight_dgrad16(fmt, w.data_ptr(), out.data_ptr(), Cin, cin_pad, Cout, _stream()), "conv3x3_weight_dgrad16")
    return out


def fc_grad_hwc_to_chw(g: torch.Tensor, C: int, HW: int) -> torch.Tensor:
    dw = torch.empty_like(g)
    check(lib.bbbp_fc_grad_hwc_to_chw_f32(g.data_ptr(), dw.data_ptr(), g.shape[0], C, HW, _stream()), "fc_grad_hwc_to_chw")
    return dw


# ---- attention --------------------------------------------------------------------------------------------------------
def attention_fwd(qkv, groups, seq, heads, head_dim, dropout_p=0.0, seed=0, want_lse=True, seed_dev=None):
    E = heads * head_dim
    out = torch.empty((groups * seq, E), device=qkv.device, dtype=torch.float32)
    lse = torch.empty((groups * seq, heads), device=qkv.device, dtype=torch.float32) if want_lse else None
    check(lib.bbbp_attention_fwd_f32(qkv.data_ptr(), out.data_ptr(), _ptr(lse), groups, seq, heads, head_dim,
                                     float(dropout_p), int(seed), _ptr(seed_dev), _stream()), "attention_fwd")
    return out, lse


def attention_bwd(qkv, out, lse, dout, groups, seq, heads, head_dim, dropout_p=0.0, seed=0, seed_dev=None):
    dqkv = torch.empty_like(qkv)
    check(lib.bbbp_attention_bwd_f32(qkv.data_ptr(), out.data_ptr(), lse.data_ptr(), dout.data_ptr(), dqkv.data_ptr(),
                                     groups, seq, heads, head_dim, float(dropout_p), int(seed), _ptr(seed_dev),
                                     _stream()), "attention_bwd")
    return dqkv


def attn_softmax_fwd(scores, scale, dropout_p=0.0, seed=0, seed_dev=None, row_base=0):
    """(rows, cols) fp32 logits -> (p, p_dropped): row softmax of scale * logits and its dropped copy (p itself at p = 0)."""
    rows, cols = scores.shape
    p = torch.empty((rows, cols), device=scores.device, dtype=torch.float32)
    pd = torch.empty_like(p) if dropout_p > 0 else None
    check(lib.bbbp_attn_softmax_fwd_f32(scores.data_ptr(), scores.stride(0), p.data_ptr(), _ptr(pd), cols, rows, cols, float(scale),
                                        float(dropout_p), int(seed), _ptr(seed_dev), int(row_base), _stream()), "attn_softmax_fwd")
    return p, (pd if pd is not None else p)


def attn_softmax_bwd(p, dpd, scale, dropout_p=0.0, seed=0, seed_dev=None, row_base=0):
    rows, cols = p.shape
    ds = torch.empty_like(p)
    check(lib.bbbp_attn_softmax_bwd_f32(p.data_ptr(), dpd.data_ptr(), ds.data_ptr(), cols, rows, cols, float(scale), float(dropout_p),
                                        int(seed), _ptr(seed_dev), int(row_base), _stream()), "attn_softmax_bwd")
    return ds


def attention_heads_bf16(qkv16: torch.Tensor, k_offset: int, v_offset: int, groups: int, seq: int, heads: int, head_dim: int,
                         ld_out: int | None = None, fmt=FMT_BF16) -> torch.Tensor:
    """Multi-head attention (head_dim 8 or 16) on the packed bf16 in_proj output; returns (groups*seq, ld_out) bf16."""
    E = heads * head_dim
    ld_out = -(-E // 8) * 8 if ld_out is None else ld_out
    out = torch.empty((groups * seq, ld_out), device=qkv16.device, dtype=_DT16[fmt])
    if ld_out != E:
        fill_zero(out[:, E:])
    check(lib.bbbp_attention_heads16(fmt, qkv16.data_ptr(), qkv16.stride(0), k_offset, v_offset, out.data_ptr(), ld_out, groups,
                                     seq, heads, head_dim, _stream()), "attention_heads16")
    return out


# ---- norms ------------------------------------------------------------------------------------------------------------
def add_layernorm_fwd(x, res, gamma, beta, eps=1e-5, save=False, bf16_ld=0):
    rows, dim = x.shape
    y = torch.empty_like(x)
    s = torch.empty_like(x) if save else None
    mean = torch.empty((rows,), device=x.device, dtype=torch.float32) if save else None
    rstd = torch.empty((rows,), device=x.device, dtype=torch.float32) if save else None
    y16 = torch.empty((rows, bf16_ld), device=x.device, dtype=torch.bfloat16) if bf16_ld else None
    check(lib.bbbp_add_layernorm_fwd_f32(x.data_ptr(), _ptr(res), gamma.data_ptr(), beta.data_ptr(), y.data_ptr(),
                                         _ptr(s), _ptr(mean), _ptr(rstd), _ptr(y16), bf16_ld, rows, dim, float(eps),
                                         _stream()), "add_layernorm_fwd")
    return y, s, mean, rstd, y16


def layernorm_fwd_pitched(x, dim, gamma, beta, eps=1e-5, ld_y=None, bf16_ld=0, fmt=FMT_BF16):
    """LN over the first ``dim`` columns of a pitched (rows, ld) buffer -> (rows, ld_y) fp32 [+ (rows, bf16_ld) 16-bit]."""
    rows = x.shape[0]
    ld_y = dim if ld_y is None else ld_y
    y = torch.empty((rows, ld_y), device=x.device, dtype=torch.float32)
    y16 = torch.empty((rows, bf16_ld), device=x.device, dtype=_DT16[fmt]) if bf16_ld else None
    check(lib.bbbp_add_layernorm_fwd_pitched16(fmt, x.data_ptr(), x.stride(0), None, 0, gamma.data_ptr(), beta.data_ptr(),
                                               y.data_ptr(), ld_y, _ptr(y16), bf16_ld, rows, dim, float(eps), _stream()),
          "add_layernorm_fwd_pitched")
    return y, y16


def ffn_layernorm16(x16, d, w1_16, b1, w2_16, b2, residual, gamma, beta, eps=1e-5, ld_y=None, ld16=0, fmt=FMT_BF16):
    """LayerNorm(residual + relu(x16 W1^T + b1) W2^T + b2) in ONE tcgen05 kernel (the hidden activation stays on the chip).
    x16: (rows, ldx) 16-bit; w1_16: (hidden, >= d); w2_16: (d, >= hidden); residual: (rows, ld_res) fp32 pitched.
    Returns ((rows, ld_y) fp32, (rows, ld16) 16-bit | None)."""
    rows, hidden = x16.shape[0], w1_16.shape[0]
    ld_y = d if ld_y is None else ld_y
    y = torch.empty((rows, ld_y), device=x16.device, dtype=torch.float32)
    y16 = torch.empty((rows, ld16), device=x16.device, dtype=_DT16[fmt]) if ld16 else None
    t0 = KERNEL_TIMER.start("ffn")
    check(lib.bbbp_ffn_layernorm16(fmt, rows, d, hidden, x16.data_ptr(), x16.stride(0), w1_16.data_ptr(), w1_16.stride(0),
                                   b1.data_ptr(), w2_16.data_ptr(), w2_16.stride(0), b2.data_ptr(), residual.data_ptr(),
                                   residual.stride(0), gamma.data_ptr(), beta.data_ptr(), float(eps), y.data_ptr(), ld_y, _ptr(y16),
                                   ld16, _stream()), "ffn_layernorm16")
    KERNEL_TIMER.stop("ffn", t0, rows)
    return y, y16


def layernorm_bwd(dy, s, mean, rstd, gamma):
    rows, dim = s.shape
    dx = torch.empty_like(s)
    dg = torch.empty_like(gamma)
    db = torch.empty_like(gamma)
    ws = torch.empty((lib.bbbp_layernorm_bwd_workspace(rows, dim) // 4,), device=s.device, dtype=torch.float32)
    check(lib.bbbp_layernorm_bwd_f32(dy.data_ptr(), s.data_ptr(), mean.data_ptr(), rstd.data_ptr(), gamma.data_ptr(),
                                     dx.data_ptr(), dg.data_ptr(), db.data_ptr(), rows, dim, ws.data_ptr(),
                                     ws.numel() * 4, _stream()), "layernorm_bwd")
    return dx, dg, db


def batchnorm_fwd(x, gamma, beta, running_mean, running_var, training, momentum=0.1, eps=1e-5):
    rows, C = x.shape
    y = torch.empty_like(x)
    sm = torch.empty((C,), device=x.device, dtype=torch.float32) if training else None
    sr = torch.empty((C,), device=x.device, dtype=torch.float32) if training else None
    check(lib.bbbp_batchnorm_fwd_f32(x.data_ptr(), gamma.data_ptr(), beta.data_ptr(), running_mean.data_ptr(),
                                     running_var.data_ptr(), y.data_ptr(), _ptr(sm), _ptr(sr), rows, C, int(training),
                                     float(momentum), float(eps), _stream()), "batchnorm_fwd")
    return y, sm, sr


def batchnorm_bwd(dy, x, gamma, save_mean, save_rstd):
    rows, C = x.shape
    dx, dg, db = torch.empty_like(x), torch.empty_like(gamma), torch.empty_like(gamma)
    check(lib.bbbp_batchnorm_bwd_f32(dy.data_ptr(), x.data_ptr(), gamma.data_ptr(), save_mean.data_ptr(),
                                     save_rstd.data_ptr(), dx.data_ptr(), dg.data_ptr(), db.data_ptr(), rows, C,
                                     _stream()), "batchnorm_bwd")
    return dx, dg, db


def batchnorm_eval_bwd(dy, x, gamma, running_mean, running_var, eps=1e-5):
    rows, C = x.shape
    dx, dg, db = torch.empty_like(x), torch.empty_like(gamma), torch.empty_like(gamma)
    check(lib.bbbp_batchnorm_eval_bwd_f32(dy.data_ptr(), x.data_ptr(), gamma.data_ptr(), running_mean.data_ptr(),
                                          running_var.data_ptr(), dx.data_ptr(), dg.data_ptr(), db.data_ptr(), rows, C,
                                          float(eps), _stream()), "batchnorm_eval_bwd")
    return dx, dg, db


# ---- fusion blocks ----------------------------------------------------------------------------------------------------
def fusion_softmax_mix_fwd(scores, c, want_w=False):
    rows, n = scores.shape
    dim = c.shape[1]
    out = torch.empty_like(c)
    w = torch.empty_like(scores) if want_w else None
    check(lib.bbbp_fusion_softmax_mix_fwd_f32(scores.data_ptr(), c.data_ptr(), out.data_ptr(), _ptr(w), rows, n, dim,
                                              _stream()), "fusion_softmax_mix_fwd")
    return out, w


def fusion_softmax_mix_bwd(w, c, dout):
    rows, n = w.shape
    dim = c.shape[1]
    dc, ds = torch.empty_like(c), torch.empty_like(w)
    check(lib.bbbp_fusion_softmax_mix_bwd_f32(w.data_ptr(), c.data_ptr(), dout.data_ptr(), dc.data_ptr(), ds.data_ptr(),
                                              rows, n, dim, _stream()), "fusion_softmax_mix_bwd")
    return dc, ds


def softmax_rows_fwd(scores):
    rows, n = scores.shape
    w = torch.empty_like(scores)
    check(lib.bbbp_softmax_rows_fwd_f32(scores.data_ptr(), w.data_ptr(), rows, n, _stream()), "softmax_rows_fwd")
    return w


def softmax_rows_bwd(w, dw):
    rows, n = w.shape
    ds = torch.empty_like(w)
    check(lib.bbbp_softmax_rows_bwd_f32(w.data_ptr(), dw.data_ptr(), ds.data_ptr(), rows, n, _stream()),
          "softmax_rows_bwd")
    return ds


def scaled_colmean_fwd(x, scale):
    """x (rows, cols) any row pitch; scale: 1-D strided view of length rows."""
    rows, cols = x.shape
    out = torch.empty((rows, cols), device=x.device, dtype=torch.float32)
    cm = torch.empty((cols,), device=x.device, dtype=torch.float32)
    check(lib.bbbp_scaled_colmean_fwd_f32(x.data_ptr(), x.stride(0), scale.data_ptr(), scale.stride(0), out.data_ptr(),
                                          cols, cm.data_ptr(), rows, cols, _stream()), "scaled_colmean_fwd")
    return out, cm


def scaled_colmean_bwd(dout, scale, colmean):
    rows, cols = dout.shape
    dx = torch.empty((rows, cols), device=dout.device, dtype=torch.float32)
    dscale = torch.empty((rows,), device=dout.device, dtype=torch.float32)
    check(lib.bbbp_scaled_colmean_bwd_f32(dout.data_ptr(), dout.stride(0), scale.data_ptr(), scale.stride(0),
                                          colmean.data_ptr(), dx.data_ptr(), cols, dscale.data_ptr(), rows, cols,
                                          _stream()), "scaled_colmean_bwd")
    return dx, dscale


# ---- elementwise ------------------------------------------------------------------------------------------------------
def act_bwd(dy, y, act):
    rows, cols = y.shape
    dx = torch.empty((rows, cols), device=y.device, dtype=torch.float32)
    check(lib.bbbp_act_bwd_f32(dy.data_ptr(), dy.stride(0), y.data_ptr(), y.stride(0), dx.data_ptr(), cols, rows, cols,
                               _ACT[act], _stream()), "act_bwd")
    return dx


def scale_by_device_scalar(x, scalar):
    y = torch.empty_like(x)
    check(lib.bbbp_scale_by_device_scalar_f32(x.data_ptr(), scalar.data_ptr(), y.data_ptr(), x.numel(), _stream()),
          "scale_by_device_scalar")
    return y


def colsum(x):
    rows, cols = x.shape
    out = torch.empty((cols,), device=x.device, dtype=torch.float32)
    need = lib.bbbp_colsum_workspace(rows, cols)
    ws = torch.empty((need // 4,), device=x.device, dtype=torch.float32) if need else None
    check(lib.bbbp_colsum_f32(x.data_ptr(), x.stride(0), out.data_ptr(), rows, cols, _ptr(ws), need, _stream()), "colsum")
    return out


def copy2d(src, dst):
    """dst[:, :] = src[:, :] for 2-D views with unit inner stride and arbitrary row pitch."""
    rows, cols = src.shape
    assert dst.shape == src.shape and src.stride(1) == 1 and dst.stride(1) == 1
    check(lib.bbbp_copy2d_f32(src.data_ptr(), src.stride(0), dst.data_ptr(), dst.stride(0), rows, cols, _stream()),
          "copy2d")
    return dst


def gather_rows(src: torch.Tensor, idx: torch.Tensor) -> torch.Tensor:
    """dst[r] = src[idx[r]] for a 2-D float32 table and a device int64 index vector."""
    rows, cols = idx.numel(), src.shape[1]
    dst = torch.empty((rows, cols), device=src.device, dtype=torch.float32)
    check(lib.bbbp_gather_rows_f32(src.data_ptr(), idx.data_ptr(), dst.data_ptr(), rows, cols, _stream()), "gather_rows")
    return dst


def scatter(src: torch.Tensor, idx: torch.Tensor, dst: torch.Tensor) -> torch.Tensor:
    """dst[idx[i]] = src[i] (float32 vectors, device int64 unique indices)."""
    assert src.dtype == torch.float32 and dst.dtype == torch.float32 and idx.dtype == torch.int64
    assert src.is_contiguous() and dst.is_contiguous() and idx.is_contiguous() and idx.numel() == src.numel()
    check(lib.bbbp_scatter_f32(src.data_ptr(), idx.data_ptr(), dst.data_ptr(), src.numel(), _stream()), "scatter")
    return dst


def dropout(x, p, seed, offset=0, seed_dev=None):
    """``seed_dev``: optional device uint64 (int64 tensor) added to ``seed`` when the kernel runs (CUDA-graph replay)."""
    y = torch.empty_like(x)
    check(lib.bbbp_dropout_f32(x.data_ptr(), y.data_ptr(), x.numel(), float(p), int(seed), int(offset), _ptr(seed_dev),
                               _stream()), "dropout")
    return y


# ---- losses / optimiser --------------------------------------------------------------------------------------------------
def mse_loss(pred, target, want_grad=True, grad_scale=1.0):
    n = pred.numel()
    loss = torch.empty((1,), device=pred.device, dtype=torch.float32)
    dpred = torch.empty((n,), device=pred.device, dtype=torch.float32) if want_grad else None
    check(lib.bbbp_mse_loss_f32(pred.data_ptr(), target.data_ptr(), loss.data_ptr(), _ptr(dpred), n, float(grad_scale),
                                _stream()), "mse_loss")
    return loss, dpred


def bce_logits_loss(logit, target, want_grad=True, grad_scale=1.0):
    n = logit.numel()
    loss = torch.empty((1,), device=logit.device, dtype=torch.float32)
    d = torch.empty((n,), device=logit.device, dtype=torch.float32) if want_grad else None
    check(lib.bbbp_bce_logits_loss_f32(logit.data_ptr(), target.data_ptr(), loss.data_ptr(), _ptr(d), n,
                                       float(grad_scale), _stream()), "bce_logits_loss")
    return loss, d


def adamw(ptr_table, sizes, chunk_tensor, chunk_offset, ntensors, nchunks, lr, beta1, beta2, eps, weight_decay, step,
          grad_scale=1.0):
    check(lib.bbbp_adamw_f32(ptr_table.data_ptr(), sizes.data_ptr(), chunk_tensor.data_ptr(), chunk_offset.data_ptr(),
                             ntensors, nchunks, float(lr), float(beta1), float(beta2), float(eps), float(weight_decay),
                             int(step), float(grad_scale), _stream()), "adamw")


def adamw_hyper(lr, beta1, beta2, eps, weight_decay, step, grad_scale=1.0) -> list[float]:
    """The eight per-step float scalars of the update, formed by the library exactly as bbbp_adamw_f32 forms them."""
    buf = (ctypes.c_float * 8)()
    check(lib.bbbp_adamw_hyper(float(lr), float(beta1), float(beta2), float(eps), float(weight_decay), int(step),
                               float(grad_scale), ctypes.cast(buf, ctypes.c_void_p)), "adamw_hyper")
    return list(buf)


def store_small(payload: bytes, dst: torch.Tensor) -> None:
    """dst's first len(payload) <= 64 bytes = payload, sent as kernel parameters (asynchronous, stream-ordered)."""
    assert len(payload) <= 64 and dst.numel() * dst.element_size() >= len(payload)
    check(lib.bbbp_store_small(payload, len(payload), dst.data_ptr(), _stream()), "store_small")


def adamw_dev(ptr_table, sizes, chunk_tensor, chunk_offset, ntensors, nchunks, hyper_dev):
    """AdamW with the per-step scalars read from device memory at run time (graph-replayable)."""
    check(lib.bbbp_adamw_dev_f32(ptr_table.data_ptr(), sizes.data_ptr(), chunk_tensor.data_ptr(), chunk_offset.data_ptr(),
                                 ntensors, nchunks, hyper_dev.data_ptr(), _stream()), "adamw_dev")


# ---- input contracts ------------------------------------------------------------------------------------------------------
def unpack_zscore(packed: torch.Tensor, n_bits: int) -> torch.Tensor:
    rows, bpr = packed.shape
    assert packed.dtype == torch.uint8 and packed.is_contiguous()
    out = torch.empty((rows, n_bits), device=packed.device, dtype=torch.float32)
    check(lib.bbbp_unpack_zscore_f32(packed.data_ptr(), bpr, out.data_ptr(), n_bits, rows, n_bits, _stream()),
          "unpack_zscore")
    return out


def u8_zscore(img: torch.Tensor) -> torch.Tensor:
    rows = img.shape[0]
    flat = img.reshape(rows, -1)
    assert flat.dtype == torch.uint8 and flat.is_contiguous()
    out = torch.empty(flat.shape, device=img.device, dtype=torch.float32)
    check(lib.bbbp_u8_zscore_f32(flat.data_ptr(), out.data_ptr(), rows, flat.shape[1], _stream()), "u8_zscore")
    return out
