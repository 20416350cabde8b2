"""Drop-in replacements for the reference's inline ``nn.Module`` classes.

Constructor signatures, ``forward(fingerprint, image)`` and every ``state_dict`` key equal the
reference's, so ``load_state_dict(torch.load("best_nn_model_maccs.pth"))``, stock ``optim.AdamW``,
``pickle.dump(model)`` and the ensemble scripts keep working unchanged:

  * transformer-CNN, canonical: Models/multi_input_data_regression_opt_transformer_cnn_20250113.py:48-119
    (identical class bodies in ..._transformer_cnn.py, ..._20250108.py, Descriptors/..._opt_all.py)
  * transformer-CNN, big:       Models/multi_input_data_regression_opt_transformer_cnn_opt_20250107_network.py:51-174
  * transformer-CNN, no fusion: Descriptors/multi_input_data_regression_opt_round_2_transformer_cnn.py:45-102
  * MLP family:                 Models/..._transformer_cnn_opt.py:52-105, ..._opt_more.py:57-110,
                                ..._rdkit.py:53-102, ..._morgan.py (== _opt)

The stock ``torch.nn`` sub-modules are kept ONLY as parameter containers (same construction order =>
same random initial weights under the same seed, same key names).  Their ``forward`` is never called:
every arithmetic step goes through the sm_100a kernels behind include/bbbp_b200.h.  There is no CPU
path; calling a model on CPU tensors raises.
"""
from __future__ import annotations

import os
import warnings

import torch
import torch.nn as nn

from . import autograd as ag

_DEFAULT_PRECISION = os.environ.get("BBBP_PRECISION", "fp32")


PRECISIONS = ("fp32", "bf16", "fp16", "strict")


def encoder_heads(fingerprint_size: int, start: int | None = None) -> int:
    """20250113.py:71-73 (start = max(1, F // 8)); 20250107_network.py:112-117 (start = 8)."""
    nhead = max(1, fingerprint_size // 8) if start is None else start
    while fingerprint_size % nhead != 0 and nhead > 1:
        nhead -= 1
    return nhead


def _stack_rows(parts):
    """cat(parts, dim=0) of 2-D float32 tensors through the library's pitched copy kernel (no ATen kernels on the path)."""
    if any(p.requires_grad for p in parts):
        return torch.cat(parts, dim=0)                 # autograd needs the graph edge (training of the big variant only)
    from . import ops
    out = torch.empty((sum(p.shape[0] for p in parts), parts[0].shape[1]), device=parts[0].device, dtype=torch.float32)
    row = 0
    for p in parts:
        ops.copy2d(p if p.stride(1) == 1 else p.contiguous(), out[row:row + p.shape[0]])
        row += p.shape[0]
    return out


def _copy_scores(dst: torch.Tensor, src: torch.Tensor) -> None:
    """dst[:] = src for two contiguous float32 score vectors, through the library's copy kernel (no ATen kernel on the path)."""
    from . import ops
    ops.copy2d(src.reshape(1, -1), dst.reshape(1, -1))


def _encoder(fingerprint_size: int, nhead: int, layers: int) -> nn.TransformerEncoder:
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")  # "enable_nested_tensor is True, but ... batch_first was not True"
        return nn.TransformerEncoder(nn.TransformerEncoderLayer(d_model=fingerprint_size, nhead=nhead), num_layers=layers)


def _score_mlp(in_dim: int, hidden: int, out_dim: int) -> nn.Sequential:
    return nn.Sequential(nn.Linear(in_dim, hidden), nn.Tanh(), nn.Linear(hidden, out_dim))


class _KernelModule(nn.Module):
    """Shared helpers: every op dispatches to the CUDA library."""

    # "fp32"   CUDA-core FMA kernels (training / validation path)
    # "bf16"   tcgen05, bf16 operands, one pass (fastest; 8-bit operand mantissa)
    # "fp16"   tcgen05, fp16 operands, one pass (same speed; 11-bit mantissa = TF32-class operands)
    # "strict" tcgen05, fp16 operands with hi + lo activation pairs through the image branch and fully split small GEMMs:
    #          |d logBB| <= 1e-3 against the fp32 reference at trained output scale
    precision = _DEFAULT_PRECISION

    # Run-time caches that must not travel with the module: CUDA graphs, streams and events cannot be pickled, and a
    # copy must not share another instance's captured buffers.  The reference pickles the whole model right after its
    # eval loop (20250113.py:229-244), i.e. AFTER graphs have been captured, and deep-copies work the same way
    # (copy.deepcopy goes through __reduce_ex__ -> __getstate__).
    _RUNTIME_STATE = ("_graphs", "_chunk_graphs", "_host_pipe", "_side_stream", "_sig_tensors", "_stack_cache")

    def __getstate__(self):
        state = dict(self.__dict__)
        for k in self._RUNTIME_STATE:
            state.pop(k, None)
        return state

    def set_precision(self, precision: str):
        if precision not in PRECISIONS:
            raise ValueError(f"precision {precision!r}: one of {PRECISIONS}")
        if getattr(self, "kind", None) == "big" and precision in ("fp16", "strict"):
            raise ValueError("the big variant (20250107) is built for the fp32 and bf16 modes")
        for m in self.modules():
            if isinstance(m, _KernelModule):
                m.precision = precision
        return self

    def _lin(self, x, layer: nn.Linear, act=None):
        return ag.linear(x, layer.weight, layer.bias, act, self.precision)

    def _bn(self, x, bn: nn.BatchNorm1d):
        training = self.training and bn.training
        use_batch = training or bn.running_mean is None
        if use_batch and x.shape[0] == 1:
            raise ValueError(f"Expected more than 1 value per channel when training, got input size {tuple(x.shape)}")
        if training and bn.track_running_stats:
            bn.num_batches_tracked += 1
        return ag.BatchNorm.apply(x, bn.weight, bn.bias, bn.running_mean, bn.running_var, use_batch,
                                  bn.momentum if bn.momentum is not None else 0.1, bn.eps)

    def _drop(self, x, module_or_p):
        p = module_or_p.p if isinstance(module_or_p, nn.Dropout) else float(module_or_p)
        return ag.dropout(x, p, self.training)

    @staticmethod
    def _check_inputs(*tensors):
        for t in tensors:
            if not t.is_cuda:
                raise RuntimeError("bbbp_b200 models run on CUDA (sm_100a) tensors only: there is no CPU fallback; "
                                   "move the model and its inputs to the GPU")
            if t.dtype != torch.float32:
                raise TypeError(f"bbbp_b200 models take float32 inputs, got {t.dtype}")


# ---- fusion blocks ---------------------------------------------------------------------------------------------------
class MultiHeadAttentionFusion(_KernelModule):
    """20250113.py:48-65: softmax over the head axis; the weights multiply the same concatenated vector."""

    def __init__(self, input_dim, num_heads=4, hidden_dim=128):
        super().__init__()
        self.attention_heads = nn.ModuleList(_score_mlp(input_dim, hidden_dim, 1) for _ in range(num_heads))
        self.softmax = nn.Softmax(dim=1)

    def _stacked_heads(self):
        """The heads' first layers stacked into one (heads*hidden, in) GEMM operand and their second layers into one
        block-diagonal (heads, heads*hidden) operand (exact: the off-diagonal zeros contribute nothing), both bf16,
        cached until any of the 4*heads tensors changes."""
        from . import ops
        fmt, split = ag.TENSOR_CORE[self.precision]
        tensors = [t for head in self.attention_heads for t in (head[0].weight, head[0].bias, head[2].weight, head[2].bias)]
        key = tuple((t.data_ptr(), t._version) for t in tensors) + (ag._weight_epoch, fmt, split)
        hit = getattr(self, "_stack_cache", None)
        if hit is not None and hit[0] == key:
            return hit[1]
        with torch.no_grad():
            nh, hid = len(self.attention_heads), self.attention_heads[0][0].out_features
            w1 = torch.cat([h[0].weight for h in self.attention_heads], dim=0)          # (nh*hid, in)   (layout only)
            b1 = torch.cat([h[0].bias for h in self.attention_heads], dim=0).contiguous()
            w2 = ops.fill_zero(torch.empty((nh, nh * hid), device=w1.device, dtype=torch.float32))
            for i, h in enumerate(self.attention_heads):
                ops.copy2d(h[2].weight.reshape(1, hid), w2[i:i + 1, i * hid:(i + 1) * hid])
            b2 = torch.cat([h[2].bias for h in self.attention_heads], dim=0).contiguous()
            val = (ops.cast16(w1.contiguous(), fmt, want_lo=split), b1, ops.cast16(w2, fmt, want_lo=split), b2, nh, hid)
        self._stack_cache = (key, val)
        return val

    def forward(self, x1, x2):
        both = ag.concat_cols(x1, x2)
        if self.precision in ag.TENSOR_CORE and not torch.is_grad_enabled():
            # inference: 2 tensor-core GEMMs for all heads instead of 2 per head (both operands split in the strict mode)
            from . import ops
            fmt, split = ag.TENSOR_CORE[self.precision]
            (w1, w1_lo), b1, (w2, w2_lo), b2, nh, hid = self._stacked_heads()
            a_hi, a_lo = ops.cast16(both, fmt, want_lo=split)
            _, h16, h16_lo = ops.gemm_bf16(a_hi, both.shape[1], w1, nh * hid, bias=b1, act="tanh", out_f32=False, out_bf16=True,
                                           fmt=fmt, a_lo=a_lo, w_lo=w1_lo, out16_lo=split)    # h16_lo is None unless split
            scores, _ = ops.gemm_bf16(h16, nh * hid, w2, nh, bias=b2, fmt=fmt, a_lo=h16_lo, w_lo=w2_lo)
            out, _ = ops.fusion_softmax_mix_fwd(scores, both)
            return out
        scores = [self._lin(self._lin(both, head[0], "tanh"), head[2]) for head in self.attention_heads]
        return ag.FusionMix.apply(ag.concat_cols(*scores), both)


class AttentionFusion(_KernelModule):
    """_rdkit.py:53-66: Softmax(dim=1) over a width-1 score, i.e. a weight that is identically 1."""

    def __init__(self, input_dim):
        super().__init__()
        self.attention = nn.Sequential(nn.Linear(input_dim, 128), nn.Tanh(), nn.Linear(128, 1), nn.Softmax(dim=1))

    def forward(self, x1, x2):
        both = ag.concat_cols(x1, x2)
        score = self._lin(self._lin(both, self.attention[0], "tanh"), self.attention[2])
        return ag.FusionMix.apply(score, both)


class MultiModalAttentionFusion(_KernelModule):
    """20250107_network.py:51-105.  The (B,1,1)*(B,D) broadcast + mean(dim=1) makes each weighted block
    ``w[i] * mean_over_the_batch(feature)``; ``groups`` > 1 evaluates several reference batches at once."""

    def __init__(self, fingerprint_dim, image_dim, hidden_dim=128):
        super().__init__()
        self.fingerprint_attention = _score_mlp(fingerprint_dim, hidden_dim, 1)
        self.image_attention = _score_mlp(image_dim, hidden_dim, 1)
        self.cross_modal_attention = _score_mlp(fingerprint_dim + image_dim, hidden_dim, fingerprint_dim)
        self.softmax = nn.Softmax(dim=1)

    def _mlp(self, x, seq_mod):
        return self._lin(self._lin(x, seq_mod[0], "tanh"), seq_mod[2])

    def forward(self, fingerprint, image, groups: int = 1):
        w_fp = self._mlp(fingerprint, self.fingerprint_attention)
        w_im = self._mlp(image, self.image_attention)
        cross = self._mlp(ag.concat_cols(fingerprint, image), self.cross_modal_attention)
        w = ag.SoftmaxRows.apply(ag.concat_cols(w_fp, w_im))
        seq = fingerprint.shape[0] // groups
        parts = []
        for g in range(groups):
            rows = slice(g * seq, (g + 1) * seq)
            fp_w = ag.ScaledColmean.apply(fingerprint[rows], w[rows, 0])
            im_w = ag.ScaledColmean.apply(image[rows], w[rows, 1])
            parts.append(ag.concat_cols(fp_w, im_w, cross[rows]))
        return parts[0] if groups == 1 else _stack_rows(parts)


# ---- transformer-CNN family --------------------------------------------------------------------------------------------
def _conv_stack(channels, image_feature_size, fc_out, dropout):
    layers, cin = [], 3
    for cout in channels:
        layers += [nn.Conv2d(cin, cout, kernel_size=3, stride=1, padding=1), nn.ReLU(), nn.MaxPool2d(kernel_size=2, stride=2)]
        cin = cout
    side = image_feature_size // (2 ** len(channels))
    layers += [nn.Flatten(), nn.Linear(cin * side * side, fc_out), nn.ReLU()]
    if dropout:
        layers.append(nn.Dropout(dropout))
    return nn.Sequential(*layers)


class TransformerCnnModel(_KernelModule):
    """kind: "canonical" | "big" | "nofusion" (see module docstring for the reference sources)."""

    IMAGE_SIDE = 128  # the reference hard-codes image.view(-1, 3, 128, 128) (20250113.py:114)

    def __init__(self, fingerprint_size, image_feature_size, kind="canonical"):
        super().__init__()
        self.kind = kind
        big = kind == "big"
        nhead = encoder_heads(fingerprint_size, 8 if big else None)
        if fingerprint_size % nhead != 0:
            raise ValueError(f"fingerprint_size={fingerprint_size} must be divisible by nhead={nhead}.")
        self.fingerprint_transformer = _encoder(fingerprint_size, nhead, 12 if big else 6)
        width = 512 if big else 128
        fp_fc = [nn.Linear(fingerprint_size, width), nn.ReLU()]
        if big:
            fp_fc.append(nn.Dropout(0.3))
        self.fingerprint_fc = nn.Sequential(*fp_fc)
        self.image_cnn = _conv_stack((64, 128, 256) if big else (32, 64), image_feature_size, width, 0.3 if big else 0.0)
        if kind == "canonical":
            self.attention_fusion = MultiHeadAttentionFusion(256, num_heads=4)
        elif big:
            self.attention_fusion = MultiModalAttentionFusion(512, 512)
        elif kind != "nofusion":
            raise ValueError(kind)
        if big:
            self.fc = nn.Sequential(nn.Linear(1536, 1024), nn.ReLU(), nn.BatchNorm1d(1024), nn.Linear(1024, 512), nn.ReLU(),
                                    nn.Dropout(0.3), nn.Linear(512, 256), nn.ReLU(), nn.Linear(256, 128), nn.ReLU(),
                                    nn.Linear(128, 64), nn.ReLU(), nn.Linear(64, 1))
        else:
            self.fc = nn.Sequential(nn.Linear(256, 256), nn.ReLU(), nn.BatchNorm1d(256), nn.Linear(256, 128), nn.ReLU(),
                                    nn.Linear(128, 64), nn.ReLU(), nn.Linear(64, 1))

    # -- encoder ------------------------------------------------------------------------------------------------------
    def _encoder_layer(self, x, layer: nn.TransformerEncoderLayer, groups: int, seq: int):
        attn = layer.self_attn
        heads = attn.num_heads
        head_dim = x.shape[1] // heads
        p_attn = float(attn.dropout) if self.training else 0.0
        qkv = ag.linear(x, attn.in_proj_weight, attn.in_proj_bias, None, self.precision)
        a = ag.attention(qkv, groups, seq, heads, head_dim, p_attn)
        sa = self._drop(self._lin(a, attn.out_proj), layer.dropout1)
        x = ag.AddLayerNorm.apply(sa, x, layer.norm1.weight, layer.norm1.bias, layer.norm1.eps)
        h = self._drop(self._lin(x, layer.linear1, "relu"), layer.dropout)
        f = self._drop(self._lin(h, layer.linear2), layer.dropout2)
        return ag.AddLayerNorm.apply(f, x, layer.norm2.weight, layer.norm2.bias, layer.norm2.eps)

    # -- encoder, inference fast path: every contraction on tcgen05, activations stay bf16 between GEMMs -----------------
    def _encoder_tensor_core_ok(self, seq: int) -> bool:
        attn = self.fingerprint_transformer.layers[0].self_attn
        head_dim = attn.embed_dim // attn.num_heads
        return (self.precision in ag.TENSOR_CORE and not self.training and (attn.num_heads == 1 or head_dim in (8, 16))
                and not (torch.is_grad_enabled() and any(p.requires_grad for p in self.fingerprint_transformer.parameters())))

    def _encoder_tensor_core(self, x, groups: int, seq: int):
        """6 x [in_proj GEMM -> (QK^T + row softmax) GEMM -> V transpose -> PV GEMM -> out_proj GEMM (+residual) -> LN
        -> FFN1 GEMM (+ReLU, bf16 out) -> FFN2 GEMM (+residual) -> LN], then fingerprint_fc.  9 launches per layer (7 for
        the many-small-heads variants, whose attention is one mma.sync flash kernel)."""
        from . import ops
        fmt, split = ag.TENSOR_CORE[self.precision]
        w16 = lambda w: ag.weight16(w, fmt)[0]
        F_ = x.shape[1]
        Fq = -(-F_ // 8) * 8                          # q | k | v each start on a 16-byte boundary
        rows = x.shape[0]
        # fp32 activations keep pitch Fq (a multiple of 4 floats) so the GEMM epilogues read residuals and write outputs
        # with 128-bit accesses even though F = 167 is prime
        if Fq != F_:
            x32 = torch.empty((rows, Fq), device=x.device, dtype=torch.float32)
            ops.copy2d(x, x32[:, :F_])
        else:
            x32 = x
        x16, _ = ops.cast16(x, fmt, ld=Fq)

        def padded_in_proj(w):                       # (3F, F) -> 16-bit (3*Fq, Fq), zero rows/cols in the pads
            out = ops.fill_zero(torch.empty((3 * Fq, Fq), device=w.device, dtype=ops._DT16[fmt]))
            for part in range(3):
                ops.cast16(w[part * F_:(part + 1) * F_], fmt, out=out[part * Fq: part * Fq + F_])
            return out

        def padded_in_bias(b):
            out = ops.fill_zero(torch.empty((1, 3 * Fq), device=b.device, dtype=torch.float32))
            for part in range(3):
                ops.copy2d(b[part * F_:(part + 1) * F_].reshape(1, F_), out[:, part * Fq: part * Fq + F_])
            return out.reshape(-1)

        for layer in self.fingerprint_transformer.layers:
            attn = layer.self_attn
            w_in = ag.derived_weight(attn.in_proj_weight, f"qkv_pad16_{fmt}", padded_in_proj)
            b_in = ag.derived_weight(attn.in_proj_bias, "qkv_pad", padded_in_bias)
            _, qkv16 = ops.gemm_bf16(x16, F_, w_in, 3 * Fq, bias=b_in, out_f32=False, out_bf16=True, fmt=fmt)
            fused_tail = False
            if attn.num_heads == 1 and seq >= self.flash_min_seq and F_ <= 192:
                # scopes wider than one score tile: streaming-softmax kernel, the seq x seq logits stay in TMEM / shared memory
                ldp = -(-seq // 8) * 8
                vt = ops.transpose_bf16(qkv16[:, 2 * Fq:], groups, seq, F_, 3 * Fq, seq * 3 * Fq, ldp)
                if self.fused_attention_tail and F_ <= 176:
                    # ... with out_proj + residual + norm1 in the kernel's tail: the attention output never reaches HBM
                    x32, x16 = ops.attention_flash_proj_ln16(qkv16, qkv16[:, Fq:], 3 * Fq, groups, seq, F_, F_ ** -0.5, vt, ldp,
                                                             w16(attn.out_proj.weight), attn.out_proj.bias, x32, layer.norm1.weight,
                                                             layer.norm1.bias, layer.norm1.eps, ld_y=Fq, ld16=Fq, fmt=fmt)
                    fused_tail = True
                else:
                    a16 = ops.attention_flash16(qkv16, qkv16[:, Fq:], 3 * Fq, groups, seq, F_, F_ ** -0.5, vt, ldp, fmt=fmt, ld_out=Fq)
            elif attn.num_heads == 1:
                p16 = ops.attention_scores_softmax_bf16(qkv16, qkv16[:, Fq:], 3 * Fq, groups, seq, F_, F_ ** -0.5, fmt=fmt)
                ldp = p16.shape[1]
                vt = ops.transpose_bf16(qkv16[:, 2 * Fq:], groups, seq, F_, 3 * Fq, seq * 3 * Fq, ldp)
                _, a16 = ops.gemm_bf16_batched(groups, seq, F_, seq, p16, ldp, seq * ldp, vt, ldp, F_ * ldp, ld_out16=Fq,
                                               fmt=fmt)
            else:       # 256 heads x 8 (2048-bit fingerprints): warp-level MMA flash kernel on the packed qkv, 16-bit out
                a16 = ops.attention_heads_bf16(qkv16, Fq, 2 * Fq, groups, seq, attn.num_heads, F_ // attn.num_heads, ld_out=Fq,
                                               fmt=fmt)
            if not fused_tail:
                s32, _ = ops.gemm_bf16(a16, F_, w16(attn.out_proj.weight), F_, bias=attn.out_proj.bias, residual=x32,
                                       ld_out=Fq, fmt=fmt)
                x32, x16 = ops.layernorm_fwd_pitched(s32, F_, layer.norm1.weight, layer.norm1.bias, layer.norm1.eps, ld_y=Fq,
                                                     bf16_ld=Fq, fmt=fmt)
            if self.fused_ffn and F_ <= 176 and layer.linear1.out_features % 128 == 0 and rows >= self.fused_ffn_min_rows:
                # linear1 + ReLU + linear2 + residual + norm2 in one kernel: the (rows, 2048) activation stays on the chip
                x32, x16 = ops.ffn_layernorm16(x16, F_, w16(layer.linear1.weight), layer.linear1.bias, w16(layer.linear2.weight),
                                               layer.linear2.bias, x32, layer.norm2.weight, layer.norm2.bias, layer.norm2.eps,
                                               ld_y=Fq, ld16=Fq, fmt=fmt)
                continue
            _, h16 = ops.gemm_bf16(x16, F_, w16(layer.linear1.weight), layer.linear1.out_features,
                                   bias=layer.linear1.bias, act="relu", out_f32=False, out_bf16=True, fmt=fmt)
            f32, _ = ops.gemm_bf16(h16, layer.linear1.out_features, w16(layer.linear2.weight), F_,
                                   bias=layer.linear2.bias, residual=x32, ld_out=Fq, fmt=fmt)
            x32, x16 = ops.layernorm_fwd_pitched(f32, F_, layer.norm2.weight, layer.norm2.bias, layer.norm2.eps, ld_y=Fq,
                                                 bf16_ld=Fq, fmt=fmt)
        fc = self.fingerprint_fc[0]
        if split:       # strict mode: fingerprint_fc with both operands split (the encoder itself is insensitive: 3e-5)
            x_hi, x_lo = ops.cast16(x32[:, :F_], fmt, ld=Fq, want_lo=True)
            w_hi, w_lo = ag.weight16(fc.weight, fmt, True)
            out, _ = ops.gemm_bf16(x_hi, F_, w_hi, fc.out_features, bias=fc.bias, act="relu", fmt=fmt, a_lo=x_lo, w_lo=w_lo)
        else:
            out, _ = ops.gemm_bf16(x16, F_, w16(fc.weight), fc.out_features, bias=fc.bias, act="relu", fmt=fmt)
        return out

    def _head(self, x):
        mods = list(self.fc)
        i = 0
        while i < len(mods):
            m = mods[i]
            if isinstance(m, nn.Linear):
                act = "relu" if i + 1 < len(mods) and isinstance(mods[i + 1], nn.ReLU) else None
                x = self._lin(x, m, act)
                i += 2 if act else 1
            elif isinstance(m, nn.BatchNorm1d):
                x = self._bn(x, m)
                i += 1
            elif isinstance(m, nn.Dropout):
                x = self._drop(x, m)
                i += 1
            else:
                raise TypeError(m)
        return x

    def _image_branch(self, image):
        side = self.IMAGE_SIDE
        mods = list(self.image_cnn)
        if (self.precision in ag.TENSOR_CORE and side == 128
                and not (torch.is_grad_enabled() and any(p.requires_grad for p in self.image_cnn.parameters()))):
            if self.kind == "big":
                x = self._image_branch_im2col(image, mods)
                return self._drop(x, mods[-1]) if isinstance(mods[-1], nn.Dropout) else x
            return self._image_branch_tensor_core(image, mods)
        if (self.precision in ag.TENSOR_CORE and side == 128 and self.kind != "big" and image.dtype == torch.float32
                and image.shape[0] >= self.tensor_core_train_min_batch and torch.is_grad_enabled()):
            # training in a tensor-core mode: forward AND backward of the branch on the tcgen05 GEMM (autograd.ImageBranchTensorCore)
            fmt, _ = ag.TENSOR_CORE[self.precision]
            return ag.ImageBranchTensorCore.apply(image, mods[0].weight, mods[0].bias, mods[3].weight, mods[3].bias,
                                                  mods[7].weight, mods[7].bias, fmt)
        x = image.reshape(-1, 3, side, side)
        i = 0
        while isinstance(mods[i], nn.Conv2d):
            x = ag.ConvReluPool.apply(x, mods[i].weight, mods[i].bias)
            i += 3  # Conv2d, ReLU, MaxPool2d
        x = x.reshape(x.shape[0], -1)          # nn.Flatten: (C, H, W) order, a view
        x = self._lin(x, mods[i + 1], "relu")
        if len(mods) > i + 3:
            x = self._drop(x, mods[i + 3])
        return x

    tensor_core_chunk = 0   # images per pass of the tcgen05 image branch (0 = all at once)
    # strict mode: background-referenced activations (False: (hi, lo) pairs in every layer), with the first layer's shifted
    # input staged as a (hi, lo) pair (two passes); the environment switches exist for the measurements in DESIGN.md
    strict_background = os.environ.get("BBBP_STRICT_BACKGROUND", "1") != "0"
    strict_conv1_split = os.environ.get("BBBP_STRICT_CONV1_SPLIT", "1") != "0"
    strict_u8_exact = os.environ.get("BBBP_STRICT_U8_EXACT", "1") != "0"
    # shortest attention scope that takes the streaming-softmax kernel (one launch instead of scores GEMM + P V GEMM).  Measured at
    # the reference's batch 256 (16 384 molecules per step): 7.23 -> 7.17 ms strict, 5.95 -> 5.72 ms bf16; 257 restores the
    # two-GEMM route for scopes that fit one score tile
    flash_min_seq = int(os.environ.get("BBBP_FLASH_MIN_SEQ", "129"))
    fused_attention_tail = os.environ.get("BBBP_FUSED_ATTENTION_TAIL", "1") != "0"   # out_proj + residual + norm1 in the flash kernel
    fused_ffn = os.environ.get("BBBP_FUSED_FFN", "1") != "0"     # encoder feed-forward + norm2 as one kernel (widths <= 176)
    # ... from this many rows up.  The fused kernel walks the 16 hidden blocks of a row tile serially (~42 us whatever the row
    # count), yet a reference-sized call is still faster with it than with two GEMMs + LayerNorm (graph-replayed call at batch
    # 32: 0.493 vs 0.525 ms, batch 256: 0.614 vs 0.657 ms with a threshold of 1 024), so the default is 0
    fused_ffn_min_rows = int(os.environ.get("BBBP_FUSED_FFN_MIN_ROWS", "0"))
    tensor_core_train_min_batch = 64   # below this the training step is launch-latency-bound and keeps the fp32 kernels
    implicit_conv = os.environ.get("BBBP_IMPLICIT_CONV", "1") != "0"    # big variant: no im2col matrix for the 64 / 128-channel layers
    implicit_chunk = 1024   # images per pass when no im2col matrix is built (bounds the NHWC activations: 1.5 MB per image)
    im2col_chunk = 256      # images per pass of the im2col route (bounds the im2col buffer: 4.7 MB per image at 64 -> 128)

    def _image_branch_im2col(self, image, mods):
        """Inference path of conv stacks the implicit-GEMM kernel is not instantiated for (the big variant's 3 -> 64 ->
        128 -> 256, 20250107_network.py:133-141): per layer bf16 im2col -> tcgen05 GEMM with bias + ReLU in the epilogue
        (NHWC bf16 out) -> 2x2 max-pool; the last activation IS the fc's K-major A operand (weight re-laid to (H,W,C))."""
        from . import ops
        convs = [m for m in mods if isinstance(m, nn.Conv2d)]
        fc = next(m for m in mods if isinstance(m, nn.Linear))
        side = self.IMAGE_SIDE
        if image.dtype == torch.uint8:
            image = ops.u8_zscore(image.reshape(-1, 3 * side * side).contiguous())
        img = image if image.is_contiguous() else image.contiguous()
        n = img.numel() // (3 * side * side)
        img = img.reshape(n, 3 * side * side)
        final_side = side >> len(convs)
        wfc = ag.derived_weight(fc.weight, "hwc_bf16",
                                lambda w: ops.fc_weight_to_hwc_bf16(w, convs[-1].out_channels, final_side * final_side))
        outs = []
        # first block on the fused tcgen05 kernel of the canonical network (instantiated for 64 output channels): conv + ReLU +
        # pool straight from the planar image, no NHWC8 copy, no im2col, no separate pooling pass
        fused_first = self.implicit_conv and side == 128 and convs[0].in_channels == 3 and convs[0].out_channels == 64
        chunk = self.implicit_chunk if fused_first else self.im2col_chunk
        for a in range(0, n, chunk):
            part = img[a:a + chunk]
            if fused_first:
                w1 = ag.derived_weight(convs[0].weight, "conv_umma_0", lambda w: ops.conv3x3_prepare_bf16(w, ops.FMT_BF16))
                x = ops.conv1_from_image_c64(part, w1, convs[0].bias, side, side)
            else:
                x = ops.image_to_nhwc8_bf16(part, 3, side, side)              # (n, H, W, 8): channels 3..7 zero
            for conv in (convs[1:] if fused_first else convs):
                nb, H, W, C = x.shape
                w16 = ag.derived_weight(conv.weight, f"im2col_{C}", lambda w, C=C: ops.conv3x3_weight_im2col_bf16(w, C))
                if self.implicit_conv and C % 64 == 0 and 128 % W == 0 and (H * W) % 128 == 0:
                    # 64 -> 128 and 128 -> 256: implicit GEMM, the A tiles are shifted TMA boxes of the activation itself
                    y = ops.conv3x3_gemm16(x, w16, conv.bias, "relu")
                else:
                    cols = ops.im2col3x3_bf16(x)
                    _, y = ops.gemm_bf16(cols, 9 * C, w16, conv.out_channels, bias=conv.bias, act="relu", out_f32=False,
                                         out_bf16=True)
                    del cols
                x = ops.maxpool2x2_nhwc_bf16(y.view(nb, H, W, conv.out_channels))
            flat = x.view(x.shape[0], -1)
            K = flat.shape[1]
            o, _ = ops.gemm_bf16(flat, K, wfc, fc.out_features, bias=fc.bias, act="relu", split_k=ops.fixed_split_k(K))
            outs.append(o)
        return outs[0] if len(outs) == 1 else _stack_rows(outs)

    def _image_branch_tensor_core(self, image, mods):
        """Inference path: planar CHW image (fp32 standardised, or raw uint8 normalised in the producer) -> tcgen05
        conv1 -> tcgen05 conv2 (each with bias + ReLU + max-pool in the TMEM epilogue) -> tcgen05 split-K GEMM against
        the (H,W,C)-re-laid fc weight.  No activation leaves the chip in fp32 and the flatten is free (NHWC rows ARE
        the fc's K-major A operand)."""
        from . import ops
        fmt, split = ag.TENSOR_CORE[self.precision]
        conv1, conv2, fc = mods[0], mods[3], mods[7]
        w1 = ag.derived_weight(conv1.weight, f"conv_umma_{fmt}", lambda w: ops.conv3x3_prepare_bf16(w, fmt))
        w2 = ag.derived_weight(conv2.weight, f"conv_umma_{fmt}", lambda w: ops.conv3x3_prepare_bf16(w, fmt))
        wfc = ag.derived_weight(fc.weight, f"hwc_{fmt}", lambda w: ops.fc_weight_to_hwc_bf16(w, 64, 32 * 32, fmt))
        img = image if image.is_contiguous() else image.contiguous()
        n = img.numel() // (3 * 128 * 128)
        img = img.reshape(n, 3 * 128 * 128)
        chunk = self.tensor_core_chunk or n
        outs = []
        for a in range(0, n, chunk):
            part = img[a:a + chunk]
            stats = ops.u8_image_stats(part) if part.dtype == torch.uint8 else None
            if split and self.strict_background:
                # strict mode: a depiction is mostly ONE value per channel, so the rounding of a 16-bit activation is the
                # SAME at every background pixel and adds up coherently through conv1 -> conv2 -> Linear(65536, 128).  Every
                # layer therefore works on activations RELATIVE to the image's background (exactly 0 on the canvas) and
                # adds back in fp32 what the constant part contributes (per-image rows, bbbp_bg_layer): conv2 and the
                # Linear take ONE pass of fp16 operands, conv1 stages x - bg as a (hi, lo) pair (strict_conv1_split)
                bg1 = ops.image_background(part, stats)
                ws1 = ag.derived_weight(conv1.weight, "tap_sums", lambda w: ops.fc_weight_channel_sums(w, 3, 9))
                ws2 = ag.derived_weight(conv2.weight, "tap_sums", lambda w: ops.fc_weight_channel_sums(w, 32, 9))
                wsf = ag.derived_weight(fc.weight, "chan_sums", lambda w: ops.fc_weight_channel_sums(w, 64, 32 * 32))
                tab1, neg2 = ops.bg_layer(ws1, conv1.bias, bg1, fmt=fmt, want_neg16=True)      # {T1, bg2}, -bg2 as fp16
                tab2, _ = ops.bg_layer(ws2, conv2.bias, tab1[:, 1], fmt=-1)                     # {T2, bg3}
                fc_add = ops.bg_layer(wsf, None, tab2[:, 1], table=False)                       # what bg3 contributes to the fc
                # raw uint8 depictions: u - background byte is an exact fp16 integer, so ONE pass has no activation rounding
                # at all (conv_umma.cu, BG = 2); standardised fp32 planes need the (hi, lo) pair
                exact = self.strict_u8_exact and part.dtype == torch.uint8
                y1 = ops.conv1_from_image_bg(part, w1, stats, bg1, tab1, fmt=fmt, split=self.strict_conv1_split and not exact)
                y2 = ops.conv3x3_relu_pool_bg(y1, w2, neg2, tab2, 64, fmt=fmt)
                o, _ = ops.gemm_bf16(y2.view(y2.shape[0], 65536), 65536, wfc, fc.out_features, bias=fc.bias, act="relu",
                                     split_k=ops.fixed_split_k_strict(65536), fmt=fmt, pre_add=fc_add)
            elif split:
                # round 2's first strict form, kept for comparison (strict_background = False): every activation of the
                # branch is a (hi, lo) fp16 pair through the same once-rounded weights (2x the MMAs of conv2 and the Linear)
                y1, y1_lo = ops.conv1_from_image_bf16(part, w1, conv1.bias, stats, fmt=fmt, split=True)
                y2, y2_lo = ops.conv3x3_relu_pool_bf16(y1, w2, conv2.bias, 64, fmt=fmt, x_lo=y1_lo)
                o, _ = ops.gemm_bf16(y2.view(y2.shape[0], 65536), 65536, wfc, fc.out_features, bias=fc.bias, act="relu",
                                     split_k=ops.fixed_split_k_strict(65536), fmt=fmt, a_lo=y2_lo.view(y2.shape[0], 65536))
            else:
                y1 = ops.conv1_from_image_bf16(part, w1, conv1.bias, stats, fmt=fmt)
                y2 = ops.conv3x3_relu_pool_bf16(y1, w2, conv2.bias, 64, fmt=fmt)
                o, _ = ops.gemm_bf16(y2.view(y2.shape[0], 65536), 65536, wfc, fc.out_features, bias=fc.bias, act="relu",
                                     split_k=ops.fixed_split_k(65536), fmt=fmt)
            outs.append(o)
        return outs[0] if len(outs) == 1 else _stack_rows(outs)

    # Set while a CUDA graph is being captured (train.GraphedTrainStep, the inference graphs below): the conv branch is
    # forked onto a second stream so the graph keeps it parallel to the encoder chain.  Forking EAGER launches of a large
    # pass was measured and dropped (8 192 molecules: 4.03 vs 3.77 ms -- the persistent conv kernels own every SM).
    fork_image_branch = False
    _side_stream = None
    image_first_min_rows = 2048   # eager inference passes of at least this many molecules launch the image branch first

    # -- CUDA-graph replay for small inference calls -------------------------------------------------------------------
    # A reference-sized call (batch 32..256) is ~80 kernel launches of a few microseconds each: launch-bound.  In eval
    # mode under no_grad the whole forward for a given (rows, groups, dtypes, precision, weight version) is captured
    # once into a CUDA graph and replayed; inputs are copied into the graph's static buffers.
    use_cuda_graphs = True
    graph_max_rows = 1024
    _graphs = None

    _sig_tensors = None

    def _apply(self, fn, *args, **kwargs):          # .to() / .cuda() / .float() replace parameter storage
        self._sig_tensors = None
        self._graphs = None
        self._chunk_graphs = None if self._chunk_graphs is not False else False
        return super()._apply(fn, *args, **kwargs)

    def _weight_signature(self):
        """Changes whenever any parameter or buffer may have changed: torch's per-tensor version counters (in-place
        updates by load_state_dict / stock optimizers) plus the epoch the fused bbbp AdamW bumps.  The tensor list is
        cached (walking the module tree costs ~0.7 ms per call, more than a reference-sized forward)."""
        if self._sig_tensors is None:
            self._sig_tensors = list(self.parameters()) + list(self.buffers())
        return (ag._weight_epoch, self._sig_tensors[0].data_ptr()) + tuple(t._version for t in self._sig_tensors)

    def _graph_signature(self, fingerprint, image, groups):
        if self.precision == "fp32":
            # the fp32 kernels read the parameters and BatchNorm buffers in place (no derived copies), so a captured graph
            # stays valid across in-place weight updates: key it on storage identity only.  A per-epoch validation loop
            # (20250113.py:195-203) then replays one graph instead of re-capturing after every optimizer step.
            if self._sig_tensors is None:
                self._sig_tensors = list(self.parameters()) + list(self.buffers())
            weights = tuple(t.data_ptr() for t in self._sig_tensors)
        else:
            weights = self._weight_signature()
        return (fingerprint.shape[0], groups, image.dtype, tuple(image.shape), self.precision, weights,
                fingerprint.device.index)

    def _forward_graphed(self, fingerprint, image, groups):
        if self._graphs is None:
            self._graphs = {}
        key = self._graph_signature(fingerprint, image, groups)
        entry = self._graphs.get(key)
        if entry is None:
            if len(self._graphs) >= 8:
                self._graphs.pop(next(iter(self._graphs)))
            s_fp, s_img = torch.empty_like(fingerprint), torch.empty_like(image)
            s_fp.copy_(fingerprint)
            s_img.copy_(image)
            cur = torch.cuda.current_stream()
            side = torch.cuda.Stream()
            side.wait_stream(cur)
            with torch.cuda.stream(side):          # warm-up: derived weights, kernel attributes, allocator pools
                for _ in range(2):
                    self._forward_groups_eager(s_fp, s_img, groups)
            cur.wait_stream(side)
            graph = torch.cuda.CUDAGraph()
            prev_fork, self.fork_image_branch = self.fork_image_branch, True    # conv branch || encoder inside the graph
            try:
                with torch.cuda.graph(graph):
                    s_out = self._forward_groups_eager(s_fp, s_img, groups)
            finally:
                self.fork_image_branch = prev_fork
            entry = (graph, s_fp, s_img, s_out)
            self._graphs[key] = entry
        graph, s_fp, s_img, s_out = entry
        s_fp.copy_(fingerprint)
        s_img.copy_(image)
        graph.replay()
        from . import ops
        out = torch.empty_like(s_out)                  # the caller owns its result: copy out of the graph's static buffer
        return ops.copy2d(s_out, out)

    def forward_groups(self, fingerprint, image, groups: int = 1):
        """``groups`` independent reference batches of equal size stacked along dim 0 (see _forward_groups_eager)."""
        if fingerprint.is_cuda and fingerprint.device.index != torch.cuda.current_device():
            # kernels are enqueued on the CURRENT device's stream: make the tensors' device current for the call
            with torch.cuda.device(fingerprint.device):
                return self.forward_groups(fingerprint, image, groups)
        from . import ops
        if (self.use_cuda_graphs and not self.training and not torch.is_grad_enabled() and fingerprint.is_cuda
                and fingerprint.shape[0] <= self.graph_max_rows and not ops.KERNEL_TIMER.names
                and fingerprint.dtype == torch.float32 and fingerprint.is_contiguous() and image.is_contiguous()
                and fingerprint.shape[0] % groups == 0 and not torch.cuda.is_current_stream_capturing()):
            try:
                return self._forward_graphed(fingerprint, image, groups)
            except Exception as e:      # capture is an optimisation: fall back to eager launches of the same kernels
                warnings.warn(f"bbbp_b200: CUDA-graph capture disabled for this model ({e})")
                self.use_cuda_graphs = False
                self._graphs = None
        return self._forward_groups_eager(fingerprint, image, groups)

    def _forward_groups_eager(self, fingerprint, image, groups: int = 1):
        """``groups`` independent reference batches of equal size stacked along dim 0 (attention and the
        big variant's batch-mean stay inside each batch).  Train-mode BatchNorm would mix
        the groups, so groups > 1 is for eval mode."""
        self._check_inputs(fingerprint)
        if image.dtype == torch.uint8:
            # raw depictions (extension of contract P2): the tcgen05 image branch normalises them in its producer;
            # every other path gets the exact ToTensor + per-molecule z-score first
            if not image.is_cuda:
                raise RuntimeError("bbbp_b200 models run on CUDA (sm_100a) tensors only: there is no CPU fallback")
            if not (self.precision in ag.TENSOR_CORE and self.kind != "big" and not torch.is_grad_enabled()):
                from . import ops
                image = ops.u8_zscore(image.reshape(fingerprint.shape[0], -1).contiguous())
        else:
            self._check_inputs(image)
        if groups > 1 and self.training:
            raise RuntimeError("forward_groups(groups > 1) is an inference path: call model.eval() first")
        rows = fingerprint.shape[0]
        if rows % groups:
            raise ValueError(f"{rows} molecules do not split into {groups} equal reference batches")
        x = fingerprint if fingerprint.is_contiguous() else fingerprint.contiguous()
        side = None
        if self.fork_image_branch and torch.cuda.is_current_stream_capturing():
            # the conv branch does not depend on the encoder: fork it so the captured graph has two parallel branches
            cur = torch.cuda.current_stream()
            if self._side_stream is None:
                self._side_stream = torch.cuda.Stream(fingerprint.device)
            side = self._side_stream
            side.wait_stream(cur)
            with torch.cuda.stream(side):
                im = self._image_branch(image)
            im.record_stream(cur)
        elif not self.training and not torch.is_grad_enabled() and rows >= self.image_first_min_rows:
            # large eager inference pass: the encoder is ~60 launches of 8-70 us, i.e. launch-bound from Python (gaps between
            # its kernels: 1.8 ms of span for 1.3 ms of work at 16 384 molecules); queued BEHIND the two long convolution
            # kernels they cost no gaps at all
            im = self._image_branch(image)
            side = False
        if self._encoder_tensor_core_ok(rows // groups):
            fp = self._encoder_tensor_core(x, groups, rows // groups)
        else:
            for layer in self.fingerprint_transformer.layers:
                x = self._encoder_layer(x, layer, groups, rows // groups)
            fp = self._lin(x, self.fingerprint_fc[0], "relu")
        if len(self.fingerprint_fc) > 2:
            fp = self._drop(fp, self.fingerprint_fc[2])
        if side is None:
            im = self._image_branch(image)
        elif side is not False:
            cur.wait_stream(side)
        if self.kind == "nofusion":
            fused = ag.concat_cols(fp, im)
        elif self.kind == "big":
            fused = self.attention_fusion(fp, im, groups)
        else:
            fused = self.attention_fusion(fp, im)
        return self._head(fused)

    def forward(self, fingerprint, image):
        return self.forward_groups(fingerprint, image, 1)

    @torch.no_grad()
    def predict_batches_packed(self, packed_bits, image_u8, batch_size: int, max_rows_per_pass: int = 16384):
        """Screening entry point on the compact input formats (SURVEY cfg4): fingerprints as little-endian packed bits
        (B, ceil(F/8)) uint8 and depictions as raw uint8 (B, 3, 128, 128).  Unpack + per-molecule z-score and the image
        normalisation run on the device, reproducing the reference's preprocessing formulas (oracle/preprocess.py)."""
        from . import ops
        n_bits = self.fingerprint_transformer.layers[0].self_attn.embed_dim
        fp = ops.unpack_zscore(packed_bits.contiguous(), n_bits)
        return self.predict_batches(fp, image_u8, batch_size, max_rows_per_pass)

    @staticmethod
    def _pipeline_spans(n: int, chunk: int, batch_size: int):
        """Chunk schedule of the copy/compute pipeline: equal chunks of whole reference batches; the ragged tail batch
        (n mod batch_size molecules) rides on the last span.  Measured on B200 (tools/e2e_sched.py, 8 192 molecules, H2D
        alone 7.25 ms): equal 1 024-molecule chunks 8.33 ms, 2 048 -> 8.55 ms, ramped schedules (short first / last
        chunks) 8.6-8.9 ms -- a chunk's graph replay costs ~0.55 ms + 0.34 ms per 1 024 molecules (the encoder is a chain
        of ~85 dependent kernels), so chunks much shorter than 1 024 no longer hide behind their own copy."""
        full = n // batch_size * batch_size
        if isinstance(chunk, (tuple, list)):
            # a SCHEDULE of chunk lengths (the last one repeats): a short first chunk exposes little of its own copy, and --
            # where the link outruns the arithmetic, as with sparse depictions (5 KB per molecule) -- every later copy hides
            # behind the chunk before it however long it is, so the remaining chunks can be few and long
            spans, a, i = [], 0, 0
            while a < full:
                b = min(full, a + chunk[min(i, len(chunk) - 1)])
                spans.append((a, b))
                a, i = b, i + 1
            chunk = max(chunk)
        else:
            spans = [(a, min(full, a + chunk)) for a in range(0, full, chunk)]
        if n > full:
            if spans and spans[-1][1] - spans[-1][0] + (n - full) <= chunk:
                spans[-1] = (spans[-1][0], n)
            else:
                spans.append((full, n))
        return spans

    @torch.no_grad()
    def predict_from_host(self, fingerprint_host, image_host, batch_size: int, chunk_molecules: int | tuple = 1024,
                          packed: bool = False, out_host: torch.Tensor | None = None, return_device: bool = False,
                          synchronize: bool = True):
        """End-to-end scoring of HOST-resident molecules (pinned tensors recommended): the host->device copy of chunk
        c+1 runs on a second stream while chunk c is being scored, so a pass costs max(copy, compute) instead of their
        sum.  ``packed`` selects the compact formats of predict_batches_packed.  Chunks are whole reference batches, so
        scores are identical to predict_batches on the same data.  Returns the (N,) scores on the host (and, with
        ``return_device``, also the device copy, e.g. for a cross-rank gather).

        ``synchronize=True`` (default) waits for the device->host copy, so the returned host tensor is ready to read.
        ``synchronize=False`` is the streaming mode for back-to-back shards: the call only enqueues work (the staging
        slots are guarded by per-slot events that persist across calls, so the first copies of the next call overlap the
        last chunk of this one); the caller synchronises the stream before reading ``out_host`` and must not reuse one
        ``out_host`` buffer for two calls in flight."""
        assert not self.training, "call model.eval() first"
        dev = next(self.parameters()).device
        if dev.index != torch.cuda.current_device():
            with torch.cuda.device(dev):
                return self.predict_from_host(fingerprint_host, image_host, batch_size, chunk_molecules, packed, out_host,
                                              return_device, synchronize)
        n = fingerprint_host.shape[0]
        if isinstance(chunk_molecules, (tuple, list)):           # a schedule of chunk lengths, see _pipeline_spans
            schedule = tuple(max(1, int(c) // batch_size) * batch_size for c in chunk_molecules)
            chunk = max(schedule)
        else:
            chunk = max(1, chunk_molecules // batch_size) * batch_size
            schedule = chunk
        scores = torch.empty((n,), device=dev, dtype=torch.float32)
        compute = torch.cuda.current_stream(dev)
        # the copy stream and the two staging slots are created once and reused: a fresh stream per call would defeat
        # the caching allocator (per-stream pools) and turn every chunk into a cudaMalloc
        sparse = hasattr(image_host, "mask") and hasattr(image_host, "offsets")      # screening.SparseDepictions
        if sparse and not packed:
            raise ValueError("sparse depictions are uint8 images: use packed=True (packed fingerprint bits + uint8 depictions)")
        key = (chunk, tuple(fingerprint_host.shape[1:]), fingerprint_host.dtype, tuple(image_host.shape[1:]), image_host.dtype,
               dev.index, sparse)
        pipe = getattr(self, "_host_pipe", None)
        if pipe is None or pipe[0] != key:
            def make_slot():
                fp_slot = torch.empty((chunk,) + tuple(fingerprint_host.shape[1:]), device=dev, dtype=fingerprint_host.dtype)
                img_slot = torch.empty((chunk,) + tuple(image_host.shape[1:]), device=dev, dtype=image_host.dtype)
                if not sparse:
                    return (fp_slot, img_slot)
                # staging for the encoded chunk (values sized for the worst case: every pixel marked) + the decode target
                return (fp_slot, img_slot, torch.empty((chunk, 2048), device=dev, dtype=torch.uint8),
                        torch.empty((chunk * 3 * 128 * 128 + 16,), device=dev, dtype=torch.uint8),
                        torch.empty((chunk + 1,), device=dev, dtype=torch.int64))
            slots = [make_slot() for _ in range(2)]
            pipe = (key, torch.cuda.Stream(dev), slots, [None, None])
            self._host_pipe = pipe
            pipe[1].wait_stream(compute)
        _, copier, slots, freed = pipe               # freed[s]: last compute that read slot s (this call or a previous one)
        ready = [None, None]
        spans = self._pipeline_spans(n, schedule, batch_size)

        def stage(i):
            a, b = spans[i]
            slot = i % 2
            with torch.cuda.stream(copier):
                if freed[slot] is not None:
                    copier.wait_event(freed[slot])        # the compute stream is done with this slot's old contents
                slots[slot][0][: b - a].copy_(fingerprint_host[a:b], non_blocking=True)
                if sparse:
                    v0, v1 = 3 * int(image_host.offsets[a]), 3 * int(image_host.offsets[b])
                    slots[slot][2][: b - a].copy_(image_host.mask[a:b], non_blocking=True)
                    if v1 > v0:
                        slots[slot][3][: v1 - v0].copy_(image_host.values[v0:v1], non_blocking=True)
                    slots[slot][4][: b - a + 1].copy_(image_host.offsets[a:b + 1], non_blocking=True)
                else:
                    slots[slot][1][: b - a].copy_(image_host[a:b], non_blocking=True)
                ready[slot] = torch.cuda.Event()
                ready[slot].record(copier)

        if spans:
            stage(0)
        for i, (a, b) in enumerate(spans):
            if i + 1 < len(spans):
                stage(i + 1)
            slot = i % 2
            compute.wait_event(ready[slot])
            fp, img = slots[slot][0][: b - a], slots[slot][1][: b - a]
            encoded = (slots[slot][2][: b - a], slots[slot][3], slots[slot][4][: b - a + 1]) if sparse else None
            part = self._score_staged_chunk(slot, fp, img, batch_size, chunk, packed, encoded)
            _copy_scores(scores[a:b], part)
            freed[slot] = torch.cuda.Event()
            freed[slot].record(compute)
        if out_host is None:
            out_host = torch.empty((n,), dtype=torch.float32, pin_memory=True)
        out_host.copy_(scores, non_blocking=True)
        if synchronize:
            done = torch.cuda.Event()
            done.record(compute)
            done.synchronize()
        return (out_host, scores) if return_device else out_host

    _chunk_graphs = None

    def _score_staged_chunk(self, slot, fp, img, batch_size, chunk, packed, encoded=None):
        """Scores of one staged chunk.  The staging slots are persistent buffers, so the whole chunk computation (unpack +
        z-score + forward of every reference batch, conv branch forked beside the encoder) is captured ONCE per (slot,
        chunk length) into a CUDA graph that reads the slot in place -- no copy into graph-private inputs, one launch per
        chunk instead of ~90, which is what lets the short first / last chunks of the ramped schedule cost less than
        their own H2D copy."""
        score = self.predict_batches_packed if packed else self.predict_batches
        if encoded is None:
            fn = score
        else:
            from . import ops

            def fn(fp_, img_, bs_, max_rows_per_pass):      # sparse depictions: rebuild the uint8 images in the slot, then score
                ops.decode_sparse_depictions(encoded[0], encoded[1], encoded[2], out=img_)
                return score(fp_, img_, bs_, max_rows_per_pass=max_rows_per_pass)
        if not self.use_cuda_graphs or self._chunk_graphs is False or torch.cuda.is_current_stream_capturing():
            return fn(fp, img, batch_size, max_rows_per_pass=chunk)
        if self._chunk_graphs is None:
            self._chunk_graphs = {}
        key = (fp.data_ptr(), img.data_ptr(), fp.shape[0], fp.dtype, img.dtype, batch_size, chunk, packed, self.precision,
               self._weight_signature(), None if encoded is None else encoded[0].data_ptr())
        entry = self._chunk_graphs.get(key)
        if entry is None:
            if len(self._chunk_graphs) >= 24:
                self._chunk_graphs.pop(next(iter(self._chunk_graphs)))
            prev_graphs, prev_fork = self.use_cuda_graphs, self.fork_image_branch
            self.use_cuda_graphs = False                 # warm-up must run the plain launches, not per-call graphs
            try:
                cur = torch.cuda.current_stream()
                side = torch.cuda.Stream(fp.device)
                side.wait_stream(cur)
                with torch.cuda.stream(side):
                    for _ in range(2):
                        fn(fp, img, batch_size, max_rows_per_pass=chunk)
                cur.wait_stream(side)
                self.fork_image_branch = True
                graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(graph):
                    out = fn(fp, img, batch_size, max_rows_per_pass=chunk)
                entry = (graph, out, fp, img, encoded)   # fp / img / encoded keep the slot views alive
                self._chunk_graphs[key] = entry
            except Exception as e:      # capture is an optimisation: fall back to plain launches of the same kernels
                warnings.warn(f"bbbp_b200: chunk-graph capture disabled for this model ({e})")
                self._chunk_graphs = False
            finally:
                self.use_cuda_graphs, self.fork_image_branch = prev_graphs, prev_fork
            if entry is None:
                return fn(fp, img, batch_size, max_rows_per_pass=chunk)
        entry[0].replay()
        return entry[1]

    @torch.no_grad()
    def predict_batches(self, fingerprint, image, batch_size: int, max_rows_per_pass: int = 16384):
        """Batched inference with the reference's batch semantics (20250113.py:229-237): molecules
        [b*batch_size, (b+1)*batch_size) form reference batch b; returns (N,) scores.  Full batches are
        evaluated ``max_rows_per_pass`` molecules at a time, the ragged tail batch on its own."""
        assert not self.training, "call model.eval() first"
        n = fingerprint.shape[0]
        out = torch.empty((n,), device=fingerprint.device, dtype=torch.float32)
        per_pass = max(1, max_rows_per_pass // batch_size) * batch_size
        full = (n // batch_size) * batch_size
        start = 0
        while start < full:
            stop = min(full, start + per_pass)
            y = self.forward_groups(fingerprint[start:stop], image[start:stop], (stop - start) // batch_size)
            _copy_scores(out[start:stop], y)
            start = stop
        if full < n:
            _copy_scores(out[full:], self.forward_groups(fingerprint[full:], image[full:], 1))
        return out


class MixedInputModel(TransformerCnnModel):
    """The canonical network: ``MixedInputModel(fingerprint_size, image_feature_size)`` (20250113.py:68-119)."""

    def __init__(self, fingerprint_size, image_feature_size):
        super().__init__(fingerprint_size, image_feature_size, "canonical")


class MixedInputModelBig(TransformerCnnModel):
    """``MixedInputModel`` of ..._opt_20250107_network.py:109-174."""

    def __init__(self, fingerprint_size, image_feature_size):
        super().__init__(fingerprint_size, image_feature_size, "big")


class MixedInputModelNoFusion(TransformerCnnModel):
    """``MixedInputModel`` of Descriptors/..._round_2_transformer_cnn.py:45-102."""

    def __init__(self, fingerprint_size, image_feature_size):
        super().__init__(fingerprint_size, image_feature_size, "nofusion")


# ---- MLP family ----------------------------------------------------------------------------------------------------------
class MlpModel(_KernelModule):
    """kind: "opt" (also _morgan) | "more" | "rdkit"; inputs are PCA-reduced feature vectors."""

    def __init__(self, fingerprint_size, image_feature_size, kind="opt"):
        super().__init__()
        self.kind = kind
        if kind == "more":
            def branch(n_in):
                return nn.Sequential(nn.Linear(n_in, 256), nn.ReLU(), nn.BatchNorm1d(256), nn.Dropout(0.3))
            self.fingerprint_fc, self.image_fc = branch(fingerprint_size), branch(image_feature_size)
            self.attention_fusion = MultiHeadAttentionFusion(512)
            self.fc = nn.Sequential(nn.Linear(512, 256), nn.ReLU(), nn.BatchNorm1d(256), nn.Linear(256, 128), nn.ReLU(),
                                    nn.Linear(128, 1))
        elif kind in ("opt", "rdkit"):
            self.fingerprint_fc = nn.Sequential(nn.Linear(fingerprint_size, 128), nn.ReLU())
            self.image_fc = nn.Sequential(nn.Linear(image_feature_size, 128), nn.ReLU())
            self.attention_fusion = AttentionFusion(256) if kind == "rdkit" else MultiHeadAttentionFusion(256)
            self.fc = nn.Sequential(nn.Linear(256, 128), nn.ReLU(), nn.Linear(128, 64), nn.ReLU(), nn.Linear(64, 1))
        else:
            raise ValueError(kind)

    def _branch(self, x, seq_mod):
        x = self._lin(x, seq_mod[0], "relu")
        if len(seq_mod) > 2:
            x = self._drop(self._bn(x, seq_mod[2]), seq_mod[3])
        return x

    def forward(self, fingerprint, image):
        self._check_inputs(fingerprint, image)
        if fingerprint.device.index != torch.cuda.current_device():
            with torch.cuda.device(fingerprint.device):
                return self.forward(fingerprint, image)
        fused = self.attention_fusion(self._branch(fingerprint, self.fingerprint_fc), self._branch(image, self.image_fc))
        return TransformerCnnModel._head(self, fused)


class MixedInputModelMLP(MlpModel):
    """``MixedInputModel`` of Models/..._transformer_cnn_opt.py:72-105 and ..._morgan.py."""

    def __init__(self, fingerprint_size, image_feature_size):
        super().__init__(fingerprint_size, image_feature_size, "opt")


class MixedInputModelMLPMore(MlpModel):
    """``MixedInputModel`` of Models/..._transformer_cnn_opt_more.py:80-110."""

    def __init__(self, fingerprint_size, image_feature_size):
        super().__init__(fingerprint_size, image_feature_size, "more")


class MixedInputModelMLPRdkit(MlpModel):
    """``MixedInputModel`` of Models/..._transformer_cnn_rdkit.py:68-102."""

    def __init__(self, fingerprint_size, image_feature_size):
        super().__init__(fingerprint_size, image_feature_size, "rdkit")


VARIANTS = {
    "tcnn": MixedInputModel, "tcnn_first": MixedInputModel, "tcnn_20250108": MixedInputModel,
    "tcnn_big": MixedInputModelBig, "tcnn_nofusion": MixedInputModelNoFusion,
    "mlp": MixedInputModelMLP, "mlp_morgan": MixedInputModelMLP, "mlp_rdkit": MixedInputModelMLPRdkit,
    "mlp_more": MixedInputModelMLPMore,
}


def build(variant: str, fingerprint_size: int, image_feature_size: int) -> nn.Module:
    """Variant names follow the reference script each class comes from (see VARIANTS)."""
    return VARIANTS[variant](fingerprint_size, image_feature_size)


def zero_dropout(model: nn.Module) -> nn.Module:
    """Set every dropout probability of ``model`` to 0 (nn.Dropout modules and nn.MultiheadAttention's own): the
    deterministic regime the parity tests and the train-step measurements run in."""
    for m in model.modules():
        if isinstance(m, nn.Dropout):
            m.p = 0.0
        if isinstance(m, nn.MultiheadAttention):
            m.dropout = 0.0
    return model


class MSELoss(nn.Module):
    """Drop-in for ``nn.MSELoss()`` (20250113.py:143): fused loss + gradient kernel."""

    def forward(self, pred, target):
        return ag.MSELoss.apply(pred, target)


class BCEWithLogitsLoss(nn.Module):
    """Extension head for the classification config (no reference NN uses BCE)."""

    def forward(self, logit, target):
        return ag.BCEWithLogitsLoss.apply(logit, target)
