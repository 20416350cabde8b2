"""CPU emulation of the tensor-core precision modes on TRAINED weights (test infrastructure, not a test).

    python tests/precision_study.py [steps]

Trains the oracle network (oracle/nets.py) on the real-depiction fixture exactly as tests/test_trained_parity_gpu.py
does, then evaluates a functional restatement of the forward pass in which the OPERANDS of every contraction are rounded
the way a tensor-core mode would round them (fp32 accumulation throughout):

  bf16 / fp16 / tf32   one pass, both operands rounded
  X+a                  activations split  a = hi + lo  (two passes against the once-rounded weights)
  Xx3                  both operands split, hi*hi + hi*lo + lo*hi

per branch of the network.  This is the measurement the "strict" mode was designed from: at a
prediction spread of ~0.63 the single-pass modes miss 1e-3 (bf16 max 2.8e-2, fp16 = tf32 max 3.0e-3); rounding the
WEIGHTS costs 2e-4 at most, rounding the ACTIVATIONS of the image branch and of the head costs 1.5-2.2e-3 each, because a
depiction is mostly one background value whose rounding error is the same at every pixel and adds up coherently.  Hence:
fp16 operands everywhere, hi + lo activations through conv1 / conv2 / Linear(65536,128), full splits in the tiny
fingerprint_fc / fusion / head GEMMs, single pass in the encoder (2.6e-5) -> max 3.4e-4.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import nets, preprocess  # noqa: E402

BRANCHES = ["enc", "attn", "fpfc", "conv1", "conv2", "imfc", "head"]


def rnd(x, mode):
    if mode == "fp32":
        return x
    if mode == "bf16":
        return x.bfloat16().float()
    if mode == "fp16":
        return x.half().float()
    if mode == "tf32":                                   # round to nearest, 10 explicit mantissa bits
        i = x.contiguous().view(torch.int32)
        return ((i + 0x1000) & ~0x1FFF).view(torch.float32)
    raise ValueError(mode)


def split(x, base):
    hi = rnd(x, base)
    return hi, rnd(x - hi, base)


def sum_preserving_round(w, base):
    """Round the 9 taps of every 3x3 filter to ``base`` with error diffusion (tap order 0..8), so that the filter's SUM --
    its response to a locally constant input, which is what a depiction's background is -- keeps full precision.  Plain
    round-to-nearest leaves each filter a random DC error of ~0.9 ulp; the same DC error at every background pixel adds up
    coherently in the Linear(65536, 128) that follows.  Float32 arithmetic, exactly as the device's weight-prepare kernel."""
    co, ci = w.shape[:2]
    flat = w.reshape(co, ci, 9)
    out = torch.empty_like(flat)
    carry = torch.zeros(co, ci, dtype=torch.float32)
    for t in range(9):
        target = flat[:, :, t] + carry
        out[:, :, t] = rnd(target, base)
        carry = target - out[:, :, t]
    return out.reshape(w.shape)


def _contract(op, a, w, mode):
    """op(a, w) with the operand precision ``mode`` (op is bilinear: a matmul or a convolution).  A trailing "/sp" rounds
    3x3 conv weights with ``sum_preserving_round`` instead of plain round-to-nearest."""
    if mode.endswith("/sp"):
        mode = mode[:-3]
        base = mode[:4]
        assert mode.endswith("+a") or mode == base
        wr = sum_preserving_round(w, base)
        if mode.endswith("+a"):
            ah, al = split(a, base)
            return op(ah, wr) + op(al, wr)
        return op(rnd(a, base), wr)
    if mode.endswith("x3"):
        ah, al = split(a, mode[:-2])
        wh, wl = split(w, mode[:-2])
        return op(ah, wh) + (op(ah, wl) + op(al, wh))
    if mode.endswith("+a"):
        ah, al = split(a, mode[:-2])
        wr = rnd(w, mode[:-2])
        return op(ah, wr) + op(al, wr)
    if ":" in mode:                                      # "fp16:a" / "fp16:w": round one operand only
        base, which = mode.split(":")
        return op(rnd(a, base) if which == "a" else a, rnd(w, base) if which == "w" else w)
    return op(rnd(a, mode), rnd(w, mode))


def mm(a, w, mode):
    return _contract(lambda x, y: x @ y.T, a, w, mode)


def conv(x, w, b, mode):
    return _contract(lambda p, q: F.conv2d(p, q, None, padding=1), x, w, mode) + b.view(1, -1, 1, 1)


def forward(sd, fp, img, modes, nhead=1, layers=6):
    """Functional restatement of MixedInputModel.forward in eval mode (20250113.py:109-119)."""
    x, Fd = fp, fp.shape[1]
    hd = Fd // nhead
    for l in range(layers):
        p = f"fingerprint_transformer.layers.{l}."
        qkv = mm(x, sd[p + "self_attn.in_proj_weight"], modes["enc"]) + sd[p + "self_attn.in_proj_bias"]
        heads = []
        for h in range(nhead):
            q, k, v = (qkv[:, i * Fd + h * hd: i * Fd + (h + 1) * hd] for i in range(3))
            prob = torch.softmax(mm(q, k, modes["attn"]) * hd ** -0.5, dim=1)
            heads.append(mm(prob, v.T.contiguous(), modes["attn"]))
        sa = mm(torch.cat(heads, 1), sd[p + "self_attn.out_proj.weight"], modes["enc"]) + sd[p + "self_attn.out_proj.bias"]
        x = F.layer_norm(x + sa, (Fd,), sd[p + "norm1.weight"], sd[p + "norm1.bias"], 1e-5)
        h1 = torch.relu(mm(x, sd[p + "linear1.weight"], modes["enc"]) + sd[p + "linear1.bias"])
        x = F.layer_norm(x + mm(h1, sd[p + "linear2.weight"], modes["enc"]) + sd[p + "linear2.bias"], (Fd,),
                         sd[p + "norm2.weight"], sd[p + "norm2.bias"], 1e-5)
    fpf = torch.relu(mm(x, sd["fingerprint_fc.0.weight"], modes["fpfc"]) + sd["fingerprint_fc.0.bias"])
    im = img.view(-1, 3, 128, 128)
    im = F.max_pool2d(torch.relu(conv(im, sd["image_cnn.0.weight"], sd["image_cnn.0.bias"], modes["conv1"])), 2)
    im = F.max_pool2d(torch.relu(conv(im, sd["image_cnn.3.weight"], sd["image_cnn.3.bias"], modes["conv2"])), 2)
    im = torch.relu(mm(im.flatten(1), sd["image_cnn.7.weight"], modes["imfc"]) + sd["image_cnn.7.bias"])
    both, hm = torch.cat((fpf, im), 1), modes["head"]
    scores = []
    for h in range(4):
        p = f"attention_fusion.attention_heads.{h}."
        t = torch.tanh(mm(both, sd[p + "0.weight"], hm) + sd[p + "0.bias"])
        scores.append(mm(t, sd[p + "2.weight"], hm) + sd[p + "2.bias"])
    fused = (torch.softmax(torch.stack(scores, 1), 1) * both[:, None, :]).sum(1)
    y = torch.relu(mm(fused, sd["fc.0.weight"], hm) + sd["fc.0.bias"])
    y = F.batch_norm(y, sd["fc.2.running_mean"], sd["fc.2.running_var"], sd["fc.2.weight"], sd["fc.2.bias"], False, 0.1, 1e-5)
    y = torch.relu(mm(y, sd["fc.3.weight"], hm) + sd["fc.3.bias"])
    y = torch.relu(mm(y, sd["fc.5.weight"], hm) + sd["fc.5.bias"])
    return mm(y, sd["fc.7.weight"], hm) + sd["fc.7.bias"]


def run(sd, fp, img, modes, bs=256):
    with torch.no_grad():
        return torch.cat([forward(sd, fp[i:i + bs], img[i:i + bs], modes).reshape(-1) for i in range(0, fp.shape[0], bs)])


def uniform(mode):
    return {k: mode for k in BRANCHES}


STRICT = dict(uniform("fp16"), conv1="fp16x3", conv2="fp16+a/sp", imfc="fp16+a", head="fp16x3", fpfc="fp16x3")


def trained_state(steps=200):
    g = np.load(os.path.join(ROOT, "tests", "golden", "b3db_depictions_u8.npz"))
    img_u8, logbb = g["img"], g["logBB"]
    n = img_u8.shape[0]
    rng = np.random.default_rng(20250113)
    bits = (rng.random((n, 167)) < 0.25).astype(np.uint8)
    bits[:, 0] = 0
    fp = torch.from_numpy(preprocess.zscore_rows(bits))
    img = torch.from_numpy(preprocess.u8_image_zscore(img_u8))
    r7 = np.random.default_rng(7)
    s = bits.astype(np.float64) @ r7.normal(size=167)
    ink = (img_u8 < 255).reshape(n, -1).mean(1)
    raw = (s - s.mean()) / s.std() + 0.7 * (ink - ink.mean()) / ink.std()
    y = torch.from_numpy((logbb.mean() + logbb.std() * (raw - raw.mean()) / raw.std()).astype(np.float32))
    torch.manual_seed(0)
    model = nets.zero_dropout(nets.build("tcnn", 167, 128))
    model.train()
    opt = torch.optim.AdamW(model.parameters(), lr=1e-3, weight_decay=1e-5)
    gen, k = torch.Generator().manual_seed(1), 0
    while k < steps:
        perm = torch.randperm(n, generator=gen)
        for a in range(0, n - 31, 32):
            nets.train_step(model, opt, fp[perm[a:a + 32]], img[perm[a:a + 32]], y[perm[a:a + 32]])
            k += 1
            if k >= steps:
                break
    model.eval()
    return model, fp, img, y


def main():
    steps = int(sys.argv[1]) if len(sys.argv) > 1 else 200
    model, fp, img, y = trained_state(steps)
    sd = model.state_dict()
    with torch.no_grad():
        module_out = torch.cat([model(fp[i:i + 256], img[i:i + 256]).reshape(-1) for i in range(0, fp.shape[0], 256)])
    base = run(sd, fp, img, uniform("fp32"))
    spread = float(module_out.std())
    r2 = 1 - float(((module_out - y) ** 2).sum() / ((y - y.mean()) ** 2).sum())
    print(f"trained {steps} steps: prediction spread {spread:.3f}, R2 {r2:.3f}; functional restatement vs module "
          f"{float((base - module_out).abs().max()):.1e}")
    f32 = uniform("fp32")
    table = {"bf16 (one pass)": uniform("bf16"), "fp16 (one pass)": uniform("fp16"), "tf32 (one pass)": uniform("tf32"),
             "fp16 (one pass, sum-preserving conv weights)": dict(uniform("fp16"), conv1="fp16/sp", conv2="fp16/sp"),
             "strict": STRICT, "strict, conv2 weights plain RN": dict(STRICT, conv2="fp16+a"),
             "strict, conv1 activations only": dict(STRICT, conv1="fp16+a"), "bf16x3 everywhere": uniform("bf16x3"), "fp16x3 everywhere": uniform("fp16x3")}
    for br in ("conv1", "conv2", "imfc", "head", "fpfc"):
        table[f"{br} fp16:a only"] = dict(f32, **{br: "fp16:a"})
        table[f"{br} fp16:w only"] = dict(f32, **{br: "fp16:w"})
    table["encoder fp16 only"] = dict(f32, enc="fp16", attn="fp16")
    table["encoder bf16 only"] = dict(f32, enc="bf16", attn="bf16")
    for name, modes in table.items():
        e = (run(sd, fp, img, modes) - base).abs()
        print(f"{name:46s} max {float(e.max()):.2e}  mean {float(e.mean()):.2e}  max/spread {float(e.max()) / spread:.4f}")


if __name__ == "__main__":
    main()
