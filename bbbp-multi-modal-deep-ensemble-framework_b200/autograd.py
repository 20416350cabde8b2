"""torch.autograd.Function wrappers: the reference's ``loss.backward()`` (20250113.py:190) keeps working
because every op of the forward registers its hand-written backward kernels here.  The autograd ENGINE
(graph walk, .grad accumulation) is torch plumbing; every arithmetic step is a bbbp_* kernel.
"""
from __future__ import annotations

import itertools
import weakref

import torch
from torch.autograd import Function

from . import ops

_seed_counter = itertools.count(1)


def next_seed() -> int:
    """Dropout seed: torch's global seed mixed with a process-wide counter (no device sync)."""
    return (torch.initial_seed() * 0x9E3779B97F4A7C15 + next(_seed_counter) * 0xD1B54A32D192ED03) & 0xFFFFFFFFFFFFFFFF


# While a training step is being captured into a CUDA graph (train.GraphedTrainStep) this holds a device int64[1]:
# the dropout kernels add its CURRENT value to their (baked) host seed each time they run, so every replay draws new masks.
_seed_dev = None


def set_seed_tensor(t):
    global _seed_dev
    prev, _seed_dev = _seed_dev, t
    return prev


# Also set during capture: a second stream for the parameter-gradient products (dW, db).  They are leaves of the
# backward dependency chain -- nothing but the optimizer consumes them -- so forking them keeps only the dX chain on the
# captured graph's critical path.  GraphedTrainStep joins the stream before the optimizer node.
_wgrad_stream = None
# Tensors the forked products still read.  Holding a reference matters for CORRECTNESS, not only for memory: when a
# gradient has two consumers (the post-norm residual), autograd's input buffer accumulates IN PLACE into the first
# arrival once it is the sole owner -- which would overwrite a dy that the side stream is still reading.  A live
# reference makes the engine add out of place instead.  Cleared when the capture ends.
_wgrad_keepalive: list = []


def set_wgrad_stream(s):
    global _wgrad_stream
    prev, _wgrad_stream = _wgrad_stream, s
    _wgrad_keepalive.clear()
    return prev


def contig(t: torch.Tensor) -> torch.Tensor:
    """Contiguous float32 copy of a 2-D strided view through the copy2d kernel (no torch arithmetic)."""
    if t.is_contiguous():
        return t
    if t.dim() == 2 and t.stride(1) == 1:
        out = torch.empty(t.shape, device=t.device, dtype=t.dtype)
        return ops.copy2d(t, out)
    return t.contiguous()


# bf16 copies of weights for the tcgen05 GEMMs, keyed by the identity of the parameter object (a weak reference
# evicts the entry when the parameter dies, so a recycled device address can never alias a stale copy).  An entry
# is refreshed when the storage moves, torch's version counter moves (load_state_dict, stock optimizers), or a bbbp
# kernel rewrote parameters behind torch's back (fused AdamW bumps the epoch).
_weight_epoch = 0
_W16_CACHE: dict = {}


def derived_weight(w: torch.Tensor, tag: str, make):
    """Cached re-layout / down-cast of parameter ``w`` (``make(w_detached)`` builds it)."""
    key = (w.data_ptr(), w._version, _weight_epoch)
    ident = (id(w), tag)
    hit = _W16_CACHE.get(ident)
    if hit is not None and hit[1] == key and hit[0]() is w:
        return hit[2]
    val = make(w.detach())
    _W16_CACHE[ident] = (weakref.ref(w, lambda _r, ident=ident: _W16_CACHE.pop(ident, None)), key, val)
    return val


# Tensor-core precision modes -> (16-bit operand format, split passes).  "strict" carries structured activations as
# hi + lo fp16 pairs (and splits both operands of the small GEMMs): |d logBB| <= 1e-3 at trained scale.
TENSOR_CORE = {"bf16": (ops.FMT_BF16, False), "fp16": (ops.FMT_F16, False), "strict": (ops.FMT_F16, True)}


def weight16(w: torch.Tensor, fmt: int, want_lo: bool = False):
    """Cached 16-bit copy of a weight as (hi, lo | None); lo = rn(w - hi) for the split GEMMs of the strict mode."""
    return derived_weight(w, f"w16_{fmt}_{int(want_lo)}", lambda d: ops.cast16(d.reshape(d.shape[0], -1), fmt, want_lo=want_lo))


def weight_bf16(w: torch.Tensor) -> torch.Tensor:
    return weight16(w, ops.FMT_BF16)[0]


def clear_weight_cache() -> None:
    global _weight_epoch
    _weight_epoch += 1


class Linear(Function):
    """y = act(x W^T + b); precision "fp32" (CUDA-core FMA) or "bf16" (tcgen05, fp32 accumulate)."""

    @staticmethod
    def forward(ctx, x, weight, bias, act, precision, training=False):
        x = contig(x)
        M, K = x.shape
        N = weight.shape[0]
        if precision in TENSOR_CORE:
            # generic nn.Linear call sites (head, fusion heads, fingerprint_fc, MLP family): tiny GEMMs, so the strict mode
            # splits BOTH operands here (hi*hi + lo*hi + hi*lo: three MMAs per K step, fp32-class products)
            fmt, split = TENSOR_CORE[precision]
            a_hi, a_lo = ops.cast16(x, fmt, want_lo=split)
            w_hi, w_lo = weight16(weight, fmt, split)
            y, _ = ops.gemm_bf16(a_hi, K, w_hi, N, bias=bias, act=act, fmt=fmt, a_lo=a_lo, w_lo=w_lo,
                                 split_k=ops.fixed_split_k_strict(K) if split else ops.fixed_split_k(K))
        else:
            # inference keeps a K-only split (bit-identical scores however batches are grouped); a training forward
            # has no such contract and takes the latency mode (one-shot kernel at M <= 32)
            # (``training`` is sampled by linear() OUTSIDE this function: inside Function.forward grad mode is always
            # off and needs_input_grad ignores torch.no_grad())
            y = ops.gemm_f32(x, weight, trans_b=True, bias=bias, act=act,
                             split_k=0 if training else ops.fixed_split_k_f32(K))
        ctx.act = act
        ctx.has_bias = bias is not None
        ctx.save_for_backward(x, weight, y if act else None)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, weight, y = ctx.saved_tensors
        dy = contig(dy)
        dpre = ops.act_bwd(dy, y, ctx.act) if ctx.act else dy
        M, K = x.shape
        N = weight.shape[0]
        dx = dw = db = None
        if ctx.needs_input_grad[0]:
            dx = ops.gemm_f32(dpre, weight, split_k=0)
        side = _wgrad_stream if torch.cuda.is_current_stream_capturing() else None
        if side is not None:
            side.wait_stream(torch.cuda.current_stream())
            dpre.record_stream(side)
            x.record_stream(side)
            _wgrad_keepalive.append((dpre, x))
        with torch.cuda.stream(side):       # stream(None) is a no-op
            if ctx.needs_input_grad[1]:
                dw = ops.gemm_f32(dpre, x, trans_a=True, split_k=0)
            if ctx.has_bias and ctx.needs_input_grad[2]:
                db = ops.colsum(dpre)
        return dx, dw, db, None, None, None


class ConvReluPool(Function):
    """maxpool2(relu(conv3x3(x) + b)), NCHW (20250113.py:85-90)."""

    @staticmethod
    def forward(ctx, x, weight, bias):
        x = x if x.is_contiguous() else x.contiguous()
        need = x.requires_grad or weight.requires_grad
        y, arg = ops.conv3x3(x, weight, bias, pool=True, want_argmax=need)
        ctx.save_for_backward(x, weight, y, arg)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, weight, y, arg = ctx.saved_tensors
        N, Cin, H, W = x.shape
        Cout = weight.shape[0]
        dy = dy if dy.is_contiguous() else dy.contiguous()
        dpre = ops.relu_pool_bwd(dy, y, arg, H, W)
        dw, db = ops.conv3x3_wgrad(dpre, x, Cout, Cin)
        dx = None
        if ctx.needs_input_grad[0]:
            dx, _ = ops.conv3x3(dpre, ops.conv3x3_flip_weights(weight), None, pool=False)
        return dx, dw, db


class ImageBranchTensorCore(Function):
    """The whole image branch of the canonical network (20250113.py:85-93: two conv + ReLU + max-pool blocks, Flatten,
    Linear(65536, 128) + ReLU) for TRAINING in a tensor-core precision mode, forward and backward, with every contraction
    on the tcgen05 GEMM and NHWC 16-bit activations in between (see csrc/conv_train.cu):

      forward   x0 = NHWC8(image); per block: cols = im2col(x) -> a = relu(cols W^T + b) -> (y, argmax) = pool(a);
                out = relu(flat(y2) Wfc_hwc^T + bfc)
      backward  Linear: dWfc = dpre^T flat, dflat = dpre Wfc_hwc (operands read in place, MN-major);
                per block: dpre = unpool(dy, argmax, y > 0); dW = dpre^T cols (MN-major both, split-K over the pixels);
                db = column sum of the masked pooled gradient; dy_prev = im2col(dpre) Wflip^T.

    The im2col rows of the forward pass are kept for the weight gradients (0.6 GB per block at batch 256).  Gradients leave
    in fp32; operands are rounded to the mode's 16-bit format once, as in the forward pass."""

    @staticmethod
    def forward(ctx, image, w1, b1, w2, b2, wfc, bfc, fmt):
        img = image if image.is_contiguous() else image.contiguous()
        n = img.numel() // (3 * 128 * 128)
        x0 = ops.image_to_nhwc8_16(img, fmt)
        saved_blocks = []
        x = x0
        for w, b in ((w1, b1), (w2, b2)):
            nb, H, W, C = x.shape
            cout = w.shape[0]
            w16 = derived_weight(w, f"im2col16_{fmt}_{C}", lambda d, C=C: ops.conv3x3_weight_im2col16(d, C, fmt))
            cols = ops.im2col3x3_16(x)
            _, a = ops.gemm_bf16(cols, 9 * C, w16, cout, bias=b, act="relu", out_f32=False, out_bf16=True, fmt=fmt)
            y, arg = ops.maxpool2x2_argmax_nhwc16(a.view(nb, H, W, cout), fmt)
            saved_blocks.append((cols, y, arg))
            x = y
        flat = x.view(n, -1)
        K = flat.shape[1]
        c_last, hw_last = x.shape[3], x.shape[1] * x.shape[2]
        wfc16 = derived_weight(wfc, f"hwc_{fmt}", lambda d: ops.fc_weight_to_hwc_bf16(d, c_last, hw_last, fmt))
        out, _ = ops.gemm_bf16(flat, K, wfc16, wfc.shape[0], bias=bfc, act="relu", split_k=ops.fixed_split_k(K), fmt=fmt)
        ctx.fmt, ctx.dims = fmt, (c_last, hw_last)
        (c1, y1, a1), (c2, y2, a2) = saved_blocks
        ctx.save_for_backward(c1, y1, a1, c2, y2, a2, out, w1, w2, wfc)
        return out

    @staticmethod
    def backward(ctx, dout):
        c1, y1, a1, c2, y2, a2, out, w1, w2, wfc = ctx.saved_tensors
        fmt = ctx.fmt
        c_last, hw_last = ctx.dims
        n = out.shape[0]
        sms = ops.sm_count(dout.device)
        # Linear(65536, 128) + ReLU
        dpre = ops.act_bwd(contig(dout), out, "relu")
        dbfc = ops.colsum(dpre)
        dp16, _ = ops.cast16(dpre, fmt)                                        # (n, 128)
        flat = y2.view(n, -1)
        K = flat.shape[1]
        g_hwc, _ = ops.gemm16_tn(dp16, flat, wfc.shape[0], K, n, trans_a=True, trans_w=True, fmt=fmt)       # dpre^T flat
        dwfc = ops.fc_grad_hwc_to_chw(g_hwc, c_last, hw_last)
        wfc16 = derived_weight(wfc, f"hwc_{fmt}", lambda d: ops.fc_weight_to_hwc_bf16(d, c_last, hw_last, fmt))
        dy, _ = ops.gemm16_tn(dp16, wfc16, n, K, wfc.shape[0], trans_w=True, fmt=fmt)                         # dpre Wfc_hwc
        grads = []
        for cols, y, arg, w, last in ((c2, y2, a2, w2, False), (c1, y1, a1, w1, True)):
            nb, OH, OW, cout = y.shape
            cin, cpad = w.shape[1], cols.shape[1] // 9
            dp, dym = ops.unpool_relu_nhwc16(dy.view(nb, OH, OW, cout), y, arg, fmt)
            db = ops.colsum(dym.view(-1, cout))
            pix = nb * 4 * OH * OW
            split = max(1, min(2 * sms, pix // 64 // 8))
            g, _ = ops.gemm16_tn(dp.view(pix, cout), cols, cout, 9 * cpad, pix, trans_a=True, trans_w=True, split_k=split, fmt=fmt)
            grads.append((ops.conv3x3_wgrad_from_im2col(g, cin, cpad), db))
            if not last:                                                       # gradient wrt the previous block's pooled output
                wd = derived_weight(w, f"dgrad16_{fmt}", lambda d: ops.conv3x3_weight_dgrad16(d, cin, fmt))
                dcols = ops.im2col3x3_16(dp)
                dy, _ = ops.gemm_bf16(dcols, 9 * cout, wd, cin, fmt=fmt)       # (pix, cin) fp32 = NHWC gradient of y_prev
                del dcols
        (dw2, db2), (dw1, db1) = grads
        return None, dw1, db1, dw2, db2, dwfc, dbfc, None


class Attention(Function):
    """Self-attention over the molecules of each reference batch; qkv rows = groups*seq.

    Three kernels behind one contract: the one-launch short-scope kernels (seq <= 32: the reference's training batch),
    the GEMM route below for single-head scopes wider than that on the training path (batch 256: six products on the
    tiled fp32 GEMM + a row-softmax kernel, 4-5x faster than the streaming kernels at that size), and the streaming
    kernels for everything else (many heads, inference)."""

    GEMM_ROUTE_MAX_SEQ = 4096        # the seq x seq probabilities are materialised (64 MB at 4096)

    @staticmethod
    def forward(ctx, qkv, groups, seq, heads, head_dim, dropout_p, seed, training=False):
        qkv = contig(qkv)
        seed_dev = _seed_dev if dropout_p > 0 else None
        ctx.cfg = (groups, seq, heads, head_dim, dropout_p, seed, seed_dev)
        ctx.route = training and heads == 1 and 32 < seq <= Attention.GEMM_ROUTE_MAX_SEQ
        if ctx.route:
            E, scale = head_dim, head_dim ** -0.5
            out = torch.empty((groups * seq, E), device=qkv.device, dtype=torch.float32)
            saved = []
            for g in range(groups):
                rows = slice(g * seq, (g + 1) * seq)
                q, k, v = qkv[rows, :E], qkv[rows, E:2 * E], qkv[rows, 2 * E:]
                s = ops.gemm_f32(q, k, trans_b=True, split_k=0)
                p, pd = ops.attn_softmax_fwd(s, scale, dropout_p, seed, seed_dev, g * seq)
                ops.gemm_f32(pd, v, out=out[rows], split_k=0)
                saved += [p] if pd is p else [p, pd]
            ctx.save_for_backward(qkv, *saved)
            return out
        out, lse = ops.attention_fwd(qkv, groups, seq, heads, head_dim, dropout_p, seed, want_lse=qkv.requires_grad,
                                     seed_dev=seed_dev)
        ctx.save_for_backward(qkv, out, lse)
        return out

    @staticmethod
    def backward(ctx, dout):
        groups, seq, heads, head_dim, dropout_p, seed, seed_dev = ctx.cfg
        dout = contig(dout)
        if ctx.route:
            qkv, *saved = ctx.saved_tensors
            E, scale = head_dim, head_dim ** -0.5
            per = 1 if dropout_p == 0 else 2
            dqkv = torch.empty_like(qkv)
            for g in range(groups):
                rows = slice(g * seq, (g + 1) * seq)
                p, pd = saved[per * g], saved[per * g + per - 1]
                q, k, v = qkv[rows, :E], qkv[rows, E:2 * E], qkv[rows, 2 * E:]
                dpd = ops.gemm_f32(dout[rows], v, trans_b=True, split_k=0)                   # dO V^T
                ops.gemm_f32(pd, dout[rows], trans_a=True, out=dqkv[rows, 2 * E:], split_k=0)     # dV = P^T dO
                ds = ops.attn_softmax_bwd(p, dpd, scale, dropout_p, seed, seed_dev, g * seq)      # carries 1/sqrt(d)
                ops.gemm_f32(ds, k, out=dqkv[rows, :E], split_k=0)                               # dQ = dS K
                ops.gemm_f32(ds, q, trans_a=True, out=dqkv[rows, E:2 * E], split_k=0)            # dK = dS^T Q
            return (dqkv,) + (None,) * 7
        qkv, out, lse = ctx.saved_tensors
        return (ops.attention_bwd(qkv, out, lse, dout, groups, seq, heads, head_dim, dropout_p, seed, seed_dev),) + (None,) * 7


class AddLayerNorm(Function):
    """LN(x + res) * gamma + beta (post-norm residual block)."""

    @staticmethod
    def forward(ctx, x, res, gamma, beta, eps):
        x = contig(x)
        res = None if res is None else contig(res)
        save = x.requires_grad or gamma.requires_grad or (res is not None and res.requires_grad)
        y, s, mean, rstd, _ = ops.add_layernorm_fwd(x, res, gamma, beta, eps, save=save)
        ctx.has_res = res is not None
        ctx.save_for_backward(s, mean, rstd, gamma)
        return y

    @staticmethod
    def backward(ctx, dy):
        s, mean, rstd, gamma = ctx.saved_tensors
        dx, dg, db = ops.layernorm_bwd(contig(dy), s, mean, rstd, gamma)
        return dx, (dx if ctx.has_res else None), dg, db, None


class BatchNorm(Function):
    """nn.BatchNorm1d: batch statistics + running update when training, running statistics otherwise."""

    @staticmethod
    def forward(ctx, x, gamma, beta, running_mean, running_var, training, momentum, eps):
        x = contig(x)
        y, sm, sr = ops.batchnorm_fwd(x, gamma, beta, running_mean, running_var, training, momentum, eps)
        ctx.training, ctx.eps = training, eps
        if training:
            ctx.save_for_backward(x, gamma, sm, sr)
        else:
            ctx.save_for_backward(x, gamma, running_mean, running_var)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, gamma, a, b = ctx.saved_tensors
        dy = contig(dy)
        if ctx.training:
            dx, dg, db = ops.batchnorm_bwd(dy, x, gamma, a, b)
        else:
            dx, dg, db = ops.batchnorm_eval_bwd(dy, x, gamma, a, b, ctx.eps)
        return dx, dg, db, None, None, None, None, None


class FusionMix(Function):
    """out = sum_h softmax_h(scores) * c  (20250113.py:62-64)."""

    @staticmethod
    def forward(ctx, scores, c):
        scores, c = contig(scores), contig(c)
        out, w = ops.fusion_softmax_mix_fwd(scores, c, want_w=True)
        ctx.save_for_backward(w, c)
        return out

    @staticmethod
    def backward(ctx, dout):
        w, c = ctx.saved_tensors
        dc, ds = ops.fusion_softmax_mix_bwd(w, c, contig(dout))
        return ds, dc


class SoftmaxRows(Function):
    @staticmethod
    def forward(ctx, scores):
        w = ops.softmax_rows_fwd(contig(scores))
        ctx.save_for_backward(w)
        return w

    @staticmethod
    def backward(ctx, dw):
        (w,) = ctx.saved_tensors
        return ops.softmax_rows_bwd(w, contig(dw))


class ScaledColmean(Function):
    """out[r, :] = scale[r] * mean_rows(x)  (the broadcast quirk of 20250107_network.py:85-96)."""

    @staticmethod
    def forward(ctx, x, scale):
        x = contig(x)
        scale = scale.reshape(-1)
        out, cm = ops.scaled_colmean_fwd(x, scale)
        ctx.save_for_backward(scale, cm)
        ctx.scale_shape = scale.shape
        return out

    @staticmethod
    def backward(ctx, dout):
        scale, cm = ctx.saved_tensors
        dx, dscale = ops.scaled_colmean_bwd(contig(dout), scale, cm)
        return dx, dscale


class Dropout(Function):
    @staticmethod
    def forward(ctx, x, p, seed):
        ctx.p, ctx.seed, ctx.seed_dev = p, seed, _seed_dev
        return ops.dropout(x if x.is_contiguous() else x.contiguous(), p, seed, seed_dev=_seed_dev)

    @staticmethod
    def backward(ctx, dy):
        return ops.dropout(dy if dy.is_contiguous() else dy.contiguous(), ctx.p, ctx.seed, seed_dev=ctx.seed_dev), None, None


class ConcatCols(Function):
    """cat(tensors, dim=1) of 2-D float32 tensors through the pitched copy kernel."""

    @staticmethod
    def forward(ctx, *tensors):
        rows = tensors[0].shape[0]
        widths = [t.shape[1] for t in tensors]
        out = torch.empty((rows, sum(widths)), device=tensors[0].device, dtype=torch.float32)
        col = 0
        for t, w in zip(tensors, widths):
            ops.copy2d(t if t.stride(1) == 1 else t.contiguous(), out[:, col:col + w])
            col += w
        ctx.widths = widths
        return out

    @staticmethod
    def backward(ctx, dout):
        grads, col = [], 0
        for w in ctx.widths:
            grads.append(dout[:, col:col + w])
            col += w
        return tuple(grads)


class MSELoss(Function):
    """mean((pred - target)^2) with the gradient produced by the same kernel (20250113.py:143,189)."""

    @staticmethod
    def forward(ctx, pred, target):
        p = pred.reshape(-1)
        p = p if p.is_contiguous() else p.contiguous()
        t = target.reshape(-1)
        t = t if t.is_contiguous() else t.contiguous()
        loss, dpred = ops.mse_loss(p, t, want_grad=pred.requires_grad)
        ctx.shape = pred.shape
        ctx.save_for_backward(dpred)
        return loss.reshape(())

    @staticmethod
    def backward(ctx, dloss):
        (dpred,) = ctx.saved_tensors
        return ops.scale_by_device_scalar(dpred, dloss.reshape(1).contiguous()).reshape(ctx.shape), None


class BCEWithLogitsLoss(Function):
    """Extension (no reference NN uses it): oracle = F.binary_cross_entropy_with_logits."""

    @staticmethod
    def forward(ctx, logit, target):
        z = logit.reshape(-1).contiguous()
        t = target.reshape(-1).contiguous()
        loss, d = ops.bce_logits_loss(z, t, want_grad=logit.requires_grad)
        ctx.shape = logit.shape
        ctx.save_for_backward(d)
        return loss.reshape(())

    @staticmethod
    def backward(ctx, dloss):
        (d,) = ctx.saved_tensors
        return ops.scale_by_device_scalar(d, dloss.reshape(1).contiguous()).reshape(ctx.shape), None


def linear(x, weight, bias=None, act=None, precision="fp32"):
    training = torch.is_grad_enabled() and (x.requires_grad or weight.requires_grad or (bias is not None and bias.requires_grad))
    return Linear.apply(x, weight, bias, act, precision, training)


def dropout(x, p, training):
    if not training or p <= 0.0:
        return x
    return Dropout.apply(x, float(p), next_seed())


def attention(qkv, groups, seq, heads, head_dim, dropout_p=0.0):
    training = torch.is_grad_enabled() and qkv.requires_grad       # sampled here: grad mode is off inside Function.forward
    return Attention.apply(qkv, groups, seq, heads, head_dim, dropout_p, next_seed() if dropout_p > 0 else 0, training)


def concat_cols(*tensors):
    return ConcatCols.apply(*tensors)
