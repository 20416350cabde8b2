"""CPU fp32 restatement of the reference networks (TEST INFRASTRUCTURE, see oracle/__init__.py).

Every network of the hot path is restated from a declarative spec with stock ``torch.nn``
modules, which is exactly where the reference's arithmetic lives.  Parameter names equal
the reference's ``state_dict`` keys so weights move freely between the reference classes,
this oracle and the CUDA product.  Construction order equals the reference's, so under
the same ``torch.manual_seed`` the random initial weights are bit-identical to the
reference's (checked in tests/test_oracle_pinning.py).

Reference sources restated here (``/root/reference/...``):
  * transformer-CNN, canonical: Models/multi_input_data_regression_opt_transformer_cnn_20250113.py:48-119
  * transformer-CNN, big:       Models/multi_input_data_regression_opt_transformer_cnn_opt_20250107_network.py:51-174
  * transformer-CNN, no fusion: Descriptors/multi_input_data_regression_opt_round_2_transformer_cnn.py:45-102
  * MLP family:                 Models/multi_input_data_regression_opt_transformer_cnn_opt.py:52-105,
                                ..._opt_more.py:57-110, ..._rdkit.py:53-102, ..._morgan.py (same as _opt)
"""
from __future__ import annotations

import torch
import torch.nn as nn


def encoder_heads(fingerprint_size: int, start: int | None = None) -> int:
    """nhead rule.  Canonical (20250113.py:71-73): start at max(1, F // 8) and decrement
    until it divides F.  Big variant (20250107_network.py:112-117): start at 8."""
    nhead = max(1, fingerprint_size // 8) if start is None else start
    while fingerprint_size % nhead != 0 and nhead > 1:
        nhead -= 1
    return nhead


def _score_mlp(in_dim: int, hidden: int, out_dim: int) -> nn.Sequential:
    return nn.Sequential(nn.Linear(in_dim, hidden), nn.Tanh(), nn.Linear(hidden, out_dim))


class MultiHeadAttentionFusion(nn.Module):
    """20250113.py:48-65.  Softmax runs over the head axis; the weights multiply the same
    concatenated vector, so the result equals ``cat(x1, x2)`` up to rounding."""

    def __init__(self, input_dim, num_heads=4, hidden_dim=128):
        super().__init__()
        self.attention_heads = nn.ModuleList(_score_mlp(input_dim, hidden_dim, 1) for _ in range(num_heads))
        self.softmax = nn.Softmax(dim=1)

    def forward(self, x1, x2):
        both = torch.cat((x1, x2), dim=1)
        scores = torch.stack([head(both) for head in self.attention_heads], dim=1)  # (B, heads, 1)
        return (self.softmax(scores) * both[:, None, :]).sum(dim=1)


class AttentionFusion(nn.Module):
    """_rdkit.py:53-66.  Softmax(dim=1) over a width-1 column is identically 1."""

    def __init__(self, input_dim):
        super().__init__()
        self.attention = nn.Sequential(nn.Linear(input_dim, 128), nn.Tanh(), nn.Linear(128, 1), nn.Softmax(dim=1))

    def forward(self, x1, x2):
        both = torch.cat((x1, x2), dim=1)
        return self.attention(both) * both


class MultiModalAttentionFusion(nn.Module):
    """20250107_network.py:51-105.  The (B,1,1)*(B,D) broadcast followed by mean(dim=1)
    makes the two weighted blocks ``w[i] * mean_over_batch(feature)``."""

    def __init__(self, fingerprint_dim, image_dim, hidden_dim=128):
        super().__init__()
        self.fingerprint_attention = _score_mlp(fingerprint_dim, hidden_dim, 1)
        self.image_attention = _score_mlp(image_dim, hidden_dim, 1)
        self.cross_modal_attention = _score_mlp(fingerprint_dim + image_dim, hidden_dim, fingerprint_dim)
        self.softmax = nn.Softmax(dim=1)

    def forward(self, fingerprint, image):
        w_fp = self.fingerprint_attention(fingerprint).unsqueeze(1)   # (B,1,1)
        w_im = self.image_attention(image).unsqueeze(1)               # (B,1,1)
        cross = self.cross_modal_attention(torch.cat((fingerprint, image), dim=1))
        w = self.softmax(torch.cat([w_fp, w_im], dim=1))              # (B,2,1)
        fp_w = (w[:, 0:1] * fingerprint).mean(dim=1)                  # (B,1,1)*(B,D) -> (B,B,D) -> (B,D)
        im_w = (w[:, 1:2] * image).mean(dim=1)
        return torch.cat((fp_w, im_w, cross), dim=1)


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


class TransformerCnnNet(nn.Module):
    """The transformer-CNN family.  ``kind``: "canonical" | "big" | "nofusion"."""

    def __init__(self, fingerprint_size, image_feature_size, kind="canonical"):
        super().__init__()
        self.kind = kind
        big = kind == "big"
        nhead = encoder_heads(fingerprint_size, 8 if big else None)
        # default batch_first=False: the (B,1,F) input is read as seq_len=B, batch=1
        self.fingerprint_transformer = nn.TransformerEncoder(
            nn.TransformerEncoderLayer(d_model=fingerprint_size, nhead=nhead), num_layers=12 if big else 6)
        width = 512 if big else 128
        fp_fc = [nn.Linear(fingerprint_size, width), nn.ReLU()]
        if big:
            fp_fc.append(nn.Dropout(0.3))
        self.fingerprint_fc = nn.Sequential(*fp_fc)
        self.image_cnn = _conv_stack((64, 128, 256) if big else (32, 64), image_feature_size, width, 0.3 if big else 0.0)
        if kind == "canonical":
            self.attention_fusion = MultiHeadAttentionFusion(256, num_heads=4)
            self.fc = nn.Sequential(nn.Linear(256, 256), nn.ReLU(), nn.BatchNorm1d(256), nn.Linear(256, 128), nn.ReLU(),
                                    nn.Linear(128, 64), nn.ReLU(), nn.Linear(64, 1))
        elif kind == "big":
            self.attention_fusion = MultiModalAttentionFusion(512, 512)
            self.fc = nn.Sequential(nn.Linear(1536, 1024), nn.ReLU(), nn.BatchNorm1d(1024), nn.Linear(1024, 512), nn.ReLU(),
                                    nn.Dropout(0.3), nn.Linear(512, 256), nn.ReLU(), nn.Linear(256, 128), nn.ReLU(),
                                    nn.Linear(128, 64), nn.ReLU(), nn.Linear(64, 1))
        elif kind == "nofusion":
            # round_2_transformer_cnn.py:78-91: plain concatenation, then the same head shape
            self.fc = nn.Sequential(nn.Linear(256, 256), nn.ReLU(), nn.BatchNorm1d(256), nn.Linear(256, 128), nn.ReLU(),
                                    nn.Linear(128, 64), nn.ReLU(), nn.Linear(64, 1))
        else:
            raise ValueError(kind)

    def forward(self, fingerprint, image):
        tokens = self.fingerprint_transformer(fingerprint.unsqueeze(1)).squeeze(1)
        fp = self.fingerprint_fc(tokens)
        im = self.image_cnn(image.view(-1, 3, 128, 128))
        fused = torch.cat((fp, im), dim=1) if self.kind == "nofusion" else self.attention_fusion(fp, im)
        return self.fc(fused)


class MlpNet(nn.Module):
    """The PCA-feature MLP family.  ``kind``: "opt" (also _morgan) | "more" | "rdkit"."""

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
        else:
            self.fingerprint_fc = nn.Sequential(nn.Linear(fingerprint_size, 128), nn.ReLU())
            self.image_fc = nn.Sequential(nn.Linear(image_feature_size, 128), nn.ReLU())
            self.attention_fusion = AttentionFusion(256) if kind == "rdkit" else MultiHeadAttentionFusion(256)
            self.fc = nn.Sequential(nn.Linear(256, 128), nn.ReLU(), nn.Linear(128, 64), nn.ReLU(), nn.Linear(64, 1))

    def forward(self, fingerprint, image):
        return self.fc(self.attention_fusion(self.fingerprint_fc(fingerprint), self.image_fc(image)))


def build(variant: str, fingerprint_size: int, image_feature_size: int) -> nn.Module:
    """variant names follow oracle/reference_classes.SCRIPTS."""
    table = {
        "tcnn": lambda: TransformerCnnNet(fingerprint_size, image_feature_size, "canonical"),
        "tcnn_first": lambda: TransformerCnnNet(fingerprint_size, image_feature_size, "canonical"),
        "tcnn_20250108": lambda: TransformerCnnNet(fingerprint_size, image_feature_size, "canonical"),
        "tcnn_big": lambda: TransformerCnnNet(fingerprint_size, image_feature_size, "big"),
        "tcnn_nofusion": lambda: TransformerCnnNet(fingerprint_size, image_feature_size, "nofusion"),
        "mlp": lambda: MlpNet(fingerprint_size, image_feature_size, "opt"),
        "mlp_morgan": lambda: MlpNet(fingerprint_size, image_feature_size, "opt"),
        "mlp_rdkit": lambda: MlpNet(fingerprint_size, image_feature_size, "rdkit"),
        "mlp_more": lambda: MlpNet(fingerprint_size, image_feature_size, "more"),
    }
    return table[variant]()


def zero_dropout(model: nn.Module) -> nn.Module:
    """Set every dropout probability to 0 (the reference's dominant regime).
    Covers nn.Dropout modules, the functional dropout inside nn.MultiheadAttention and
    the encoder layers' dropout modules."""
    for m in model.modules():
        if isinstance(m, nn.Dropout):
            m.p = 0.0
        if isinstance(m, nn.MultiheadAttention):
            m.dropout = 0.0
    return model


def train_step(model: nn.Module, optimizer, fingerprint, image, target) -> torch.Tensor:
    """One inner-loop iteration, 20250113.py:187-191 (MSELoss on the squeezed prediction)."""
    optimizer.zero_grad()
    loss = nn.functional.mse_loss(model(fingerprint, image).squeeze(), target)
    loss.backward()
    optimizer.step()
    return loss.detach()
