"""Parity at a REALISTIC output scale.

Random-init networks answer almost the same number for every molecule (std 6e-3), so an absolute tolerance says
little there.  Here the weights are TRAINED until the predictions spread like logBB does (std ~0.6-0.75), on the
1 058 real depictions the reference ships (tests/golden/b3db_depictions_u8.npz, made by oracle/make_real_fixture.py
with the reference's own Resize + ToTensor pipeline), and every precision mode is held to a bound at that scale:

  fp32    CUDA-core kernels                         |d| <= 1e-3 (north_star), R2 / MSE / AUC equal to 3 decimals
  strict  tcgen05, fp16 operands, activations of the structured branches carried as hi + lo pairs, tiny head GEMMs
          with both operands split                  |d| <= 1e-3, metrics equal to 3 decimals
  fp16    tcgen05, fp16 operands, one pass          |d| <= FP16_REL * spread   (TF32-class: 11-bit operands)
  bf16    tcgen05, bf16 operands, one pass          |d| <= BF16_REL * spread   (8-bit operands)

Where the relative bounds come from: unit round-off u = 2^-12 (fp16, round to nearest) or 2^-9 (bf16).  Weight
rounding and the rounding of unstructured activations average out over the K = 65 536 / 288 / 2 048 long dot
products; what does NOT average is the rounding of STRUCTURED activations -- a depiction is mostly one background
value, so every background pixel carries the same rounding error through conv1 -> conv2 -> Linear(65536, 128) -> head,
four coherent stages with gains of a few units each.  A CPU emulation of the four modes on these very weights
(tests/precision_study.py, same operand rounding, fp32 accumulation) gives max |d| / spread = 0.044 (bf16), 0.0048
(fp16), 0.0005 (strict); the asserted factors are twice the emulated maxima.  MACCS bits need RDKit (absent here), so
fingerprints are seeded Bernoulli(0.25) rows with bit 0 clear, standardised exactly as the reference does.
"""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN
from oracle import nets, preprocess

pytestmark = pytest.mark.gpu
IMG = 3 * 128 * 128
BF16_REL, FP16_REL = 0.09, 0.01        # max |d| as a fraction of the spread (std) of the oracle's predictions
BF16_MEAN_REL, FP16_MEAN_REL = 0.02, 0.0025


def _real_set():
    g = np.load(os.path.join(GOLDEN, "b3db_depictions_u8.npz"))
    return g["img"], g["logBB"]


def _learnable_labels(bits, img_u8, logbb):
    """Labels with the mean / spread of the real logBB column but a learnable dependence on BOTH inputs (a fixed random
    linear read-out of the bits + the ink fraction of the depiction), so a short training run gives informative
    predictions (R2 ~ 0.5) instead of a constant."""
    rng = np.random.default_rng(7)
    s = bits.astype(np.float64) @ rng.normal(size=bits.shape[1])
    ink = (img_u8 < 255).reshape(len(img_u8), -1).mean(1)
    raw = (s - s.mean()) / s.std() + 0.7 * (ink - ink.mean()) / ink.std()
    return (logbb.mean() + logbb.std() * (raw - raw.mean()) / raw.std()).astype(np.float32)


def _metrics(pred, y):
    mse = float(((pred - y) ** 2).mean())
    r2 = 1.0 - float(((pred - y) ** 2).sum() / ((y - y.mean()) ** 2).sum())
    return mse, r2


def _oracle_scores(ref, fp, img, bs):
    with torch.no_grad():
        return torch.cat([ref(fp[i:i + bs], img[i:i + bs]).reshape(-1) for i in range(0, fp.shape[0], bs)])


def _train_on_device(model, fp, img, y, steps, lr, batch, criterion):
    """Weights for the variants that are too slow to train on the CPU: the product's own fp32 training step (itself held
    to the oracle's gradients in test_model_gpu.py).  How the weights were obtained does not matter for the comparison:
    the oracle and the product evaluate the SAME state_dict afterwards."""
    import bbbp_b200
    nets.zero_dropout(model)
    model.train().set_precision("fp32")
    opt = bbbp_b200.AdamW(model.parameters(), lr=lr, weight_decay=1e-5)
    step = bbbp_b200.GraphedTrainStep(model, opt, criterion)
    g = torch.Generator().manual_seed(3)
    n, k = fp.shape[0], 0
    while k < steps:
        perm = torch.randperm(n, generator=g)
        for a in range(0, n - batch + 1, batch):
            idx = perm[a:a + batch].cuda()
            step(fp[idx], img[idx], y[idx])
            k += 1
            if k >= steps:
                break
    model.eval()
    return model


@pytest.fixture(scope="module")
def maccs_trained(cuda_device):
    """The canonical MACCS network (20250113.py:68-119) after 200 AdamW steps of the reference loop body (batch 32, dropout
    off = the reference's regime after epoch 1; lr 1e-3 to get there in 200 steps instead of 2 000)."""
    import bbbp_b200
    img_u8, logbb = _real_set()
    n = img_u8.shape[0]
    rng = np.random.default_rng(20250113)
    bits = (rng.random((n, 167)) < 0.25).astype(np.uint8)
    bits[:, 0] = 0
    fp = torch.from_numpy(preprocess.zscore_rows(bits))
    img = torch.from_numpy(preprocess.u8_image_zscore(img_u8))
    y = torch.from_numpy(_learnable_labels(bits, img_u8, logbb))
    # Trained by the product's own fp32 step (held to the oracle's gradients in test_model_gpu.py) because that is
    # REPRODUCIBLE: its kernels are deterministic, whereas 200 chaotic lr-1e-3 steps on the CPU end somewhere else for every
    # host thread count (prediction spread 0.604 with 8 threads, 0.642 with 16; R2 +0.65 vs -0.63), which made the errors
    # below -- and whether they met their bounds -- depend on the box.  How the weights were obtained does not matter for
    # the comparison: the oracle and the product evaluate the SAME state_dict.
    torch.manual_seed(0)
    ours = bbbp_b200.build("tcnn", 167, 128).to(cuda_device)
    _train_on_device(ours, fp.cuda(), img.cuda(), y.cuda(), 200, 1e-3, 32, bbbp_b200.MSELoss())
    ref = nets.build("tcnn", 167, 128)
    ref.load_state_dict({k: v.cpu() for k, v in ours.state_dict().items()}, strict=True)
    ref.eval()
    want = _oracle_scores(ref, fp, img, 256)
    spread = float(want.std())
    assert spread > 0.3, f"training did not spread the predictions (std {spread})"
    return dict(ref=ref, ours=ours, fp=fp, img=img, y=y, want=want, spread=spread, bits=bits, img_u8=img_u8)


def _check_mode(got, want, y, spread, mode, what):
    d = (got - want).abs()
    mx, mean = float(d.max()), float(d.mean())
    print(f"[trained parity] {what} {mode}: max |d| {mx:.3e} mean {mean:.3e} spread {spread:.3f}")
    if mode in ("fp32", "strict"):
        assert mx <= 1e-3, f"{what} {mode}: max |d logBB| {mx:.3e} > 1e-3"
        (mse_g, r2_g), (mse_w, r2_w) = _metrics(got, y), _metrics(want, y)
        assert round(mse_g, 3) == round(mse_w, 3) and round(r2_g, 3) == round(r2_w, 3), (mse_g, mse_w, r2_g, r2_w)
    else:
        rel, mrel = (BF16_REL, BF16_MEAN_REL) if mode == "bf16" else (FP16_REL, FP16_MEAN_REL)
        assert mx <= rel * spread and mean <= mrel * spread, f"{what} {mode}: max {mx:.3e} mean {mean:.3e} spread {spread:.3f}"
        (mse_g, r2_g), (mse_w, r2_w) = _metrics(got, y), _metrics(want, y)
        assert abs(r2_g - r2_w) <= 4 * rel and abs(mse_g - mse_w) <= 4 * rel * spread ** 2


MODES = ["fp32", "strict", "fp16", "bf16"]


@pytest.mark.parametrize("mode", MODES)
def test_maccs_b256_trained_weights_real_depictions(maccs_trained, mode):
    """BASELINE configs[0]: 1 058 molecules, batch 256 (4 x 256 + 34), informative predictions (R2 ~ 0.5)."""
    t = maccs_trained
    assert t["spread"] > 0.3                             # the metric comparison below is not about a constant predictor
    ours = t["ours"].set_precision(mode)
    got = ours.predict_batches(t["fp"].cuda(), t["img"].cuda(), 256).cpu()
    _check_mode(got, t["want"], t["y"], t["spread"], mode, "MACCS b256 (fp32 contract)")


@pytest.mark.parametrize("mode", MODES)
def test_maccs_b256_trained_weights_packed_bits_and_uint8_depictions(maccs_trained, mode):
    """The compact input contract (packed bits + raw uint8 depictions, SURVEY cfg4) on the same trained network at the
    benched batch size: unpack + z-score and the image normalisation run on the device."""
    t = maccs_trained
    ours = t["ours"].set_precision(mode)
    packed = torch.from_numpy(preprocess.pack_bits(t["bits"])).cuda()
    got = ours.predict_batches_packed(packed, torch.from_numpy(t["img_u8"]).cuda(), 256).cpu()
    _check_mode(got, t["want"], t["y"], t["spread"], mode, "MACCS b256 (packed + uint8)")


def test_real_depictions_through_the_image_contract(cuda_device):
    """P2 on real data: the device's uint8 -> ToTensor -> per-molecule z-score equals the reference formula on all
    1 058 shipped depictions (float64 statistics, <= 1 ulp of float32 on the values)."""
    import bbbp_b200
    img_u8, _ = _real_set()
    want = preprocess.u8_image_zscore(img_u8)
    got = bbbp_b200.ops.u8_zscore(torch.from_numpy(img_u8).cuda()).cpu().numpy()
    assert got.shape == want.shape
    err = np.abs(got - want)
    assert float(err.max()) <= 4e-6 * max(1.0, float(np.abs(want).max())), float(err.max())


@pytest.fixture(scope="module")
def morgan_trained(cuda_device):
    """BASELINE configs[2]: the 2048-bit variant (256 heads x 8, 160 M parameters) as a classifier, trained for 60 steps
    of the product's fp32 BCE step at batch 32 on 512 real depictions + Bernoulli(0.022) Morgan-like bits."""
    import bbbp_b200
    img_u8, _ = _real_set()
    n = 512
    img_u8 = img_u8[:n]
    rng = np.random.default_rng(20250115)
    bits = (rng.random((n, 2048)) < 0.022).astype(np.uint8)
    fp = torch.from_numpy(preprocess.zscore_rows(bits))
    img = torch.from_numpy(preprocess.u8_image_zscore(img_u8))
    s = bits.astype(np.float64) @ rng.normal(size=2048)
    ink = (img_u8 < 255).reshape(n, -1).mean(1)
    raw = (s - s.mean()) / s.std() + 0.7 * (ink - ink.mean()) / ink.std()
    y = torch.from_numpy((raw > np.quantile(raw, 0.36)).astype(np.float32))          # 64 % positives like B3DB
    torch.manual_seed(1)
    ours = bbbp_b200.build("tcnn", 2048, 128).to(cuda_device)
    _train_on_device(ours, fp.cuda(), img.cuda(), y.cuda(), 60, 3e-4, 32, bbbp_b200.BCEWithLogitsLoss())
    ref = nets.build("tcnn", 2048, 128)
    ref.load_state_dict({k: v.cpu() for k, v in ours.state_dict().items()}, strict=True)
    ref.eval()
    want = _oracle_scores(ref, fp, img, 256)
    return dict(ours=ours, fp=fp, img=img, y=y, want=want, spread=float(want.std()))


@pytest.mark.parametrize("mode", MODES)
def test_morgan2048_b256_trained_classifier_logits_and_auc(morgan_trained, mode):
    from sklearn.metrics import roc_auc_score
    t = morgan_trained
    assert t["spread"] > 0.2, t["spread"]
    ours = t["ours"].set_precision(mode)
    got = ours.predict_batches(t["fp"].cuda(), t["img"].cuda(), 256).cpu()
    d = (got - t["want"]).abs()
    print(f"[trained parity] Morgan-2048 b256 {mode}: max |d logit| {float(d.max()):.3e} mean {float(d.mean()):.3e} "
          f"spread {t['spread']:.3f}")
    y = t["y"].numpy().astype(int)
    auc_w, auc_g = roc_auc_score(y, t["want"].numpy()), roc_auc_score(y, got.numpy())
    assert 0.55 < auc_w, auc_w                               # informative scores
    if mode in ("fp32", "strict"):
        assert float(d.max()) <= 1e-3
        assert round(auc_g, 3) == round(auc_w, 3), (auc_g, auc_w)
        p_w, p_g = torch.sigmoid(t["want"]), torch.sigmoid(got)
        assert float((p_w - p_g).abs().max()) <= 1e-3       # north_star: probabilities within 1e-3
    else:
        rel = BF16_REL if mode == "bf16" else FP16_REL
        assert float(d.max()) <= rel * t["spread"]
        assert abs(auc_g - auc_w) <= rel


@pytest.fixture(scope="module")
def big_trained(cuda_device):
    """The big variant (20250107_network.py:109-174; 12 layers, 64/128/256 convs) trained for 40 fp32 steps at batch 32."""
    import bbbp_b200
    img_u8, logbb = _real_set()
    n = 256
    img_u8 = img_u8[::4][:n]
    rng = np.random.default_rng(20250107)
    bits = (rng.random((n, 167)) < 0.25).astype(np.uint8)
    bits[:, 0] = 0
    fp = torch.from_numpy(preprocess.zscore_rows(bits))
    img = torch.from_numpy(preprocess.u8_image_zscore(img_u8))
    y = torch.from_numpy(_learnable_labels(bits, img_u8, logbb[::4][:n]))
    torch.manual_seed(2)
    ours = bbbp_b200.build("tcnn_big", 167, 128).to(cuda_device)
    _train_on_device(ours, fp.cuda(), img.cuda(), y.cuda(), 40, 3e-4, 32, bbbp_b200.MSELoss())
    ref = nets.build("tcnn_big", 167, 128)
    ref.load_state_dict({k: v.cpu() for k, v in ours.state_dict().items()}, strict=True)
    ref.eval()
    want = _oracle_scores(ref, fp, img, 32)
    return dict(ours=ours, fp=fp, img=img, y=y, want=want, spread=float(want.std()))


@pytest.mark.parametrize("mode", ["fp32", "bf16"])
def test_big_variant_b32_trained_weights(big_trained, mode):
    t = big_trained
    ours = t["ours"].set_precision(mode)
    got = ours.predict_batches(t["fp"].cuda(), t["img"].cuda(), 32).cpu()
    d = (got - t["want"]).abs()
    scale = max(t["spread"], float(t["want"].abs().mean()))
    print(f"[trained parity] big b32 {mode}: max |d| {float(d.max()):.3e} mean {float(d.mean()):.3e} spread {t['spread']:.3f}")
    if mode == "fp32":
        assert float(d.max()) <= 1e-3
    else:
        assert float(d.max()) <= BF16_REL * scale
