"""Pin the CPU oracle (oracle/nets.py, oracle/preprocess.py) before trusting it:
  * against golden vectors generated from the REFERENCE's own classes and shipped artefacts
    (tests/golden/*.npz, made by oracle/make_golden.py in the build container);
  * against the reference classes themselves when /root/reference is present (build container only).
"""
import json
import os

import numpy as np
import pytest
import torch

from conftest import load_golden, seeded_inputs, GOLDEN
from oracle import nets, preprocess, reference_classes as rc

torch.set_num_threads(max(1, min(8, os.cpu_count() or 1)))
IMG = 3 * 128 * 128


def _state_from_golden(g):
    return {k[len("param:"):]: torch.from_numpy(g[k]) for k in g.files if k.startswith("param:")}


@pytest.mark.parametrize("name,fp_dim,img_dim,kat", [
    ("mlp_ckpt_maccs", 64, 128, [-0.0565892, -0.1827354, -0.1608347, -0.1144812]),
    ("mlp_ckpt_morgan", 128, 256, None),
])
def test_shipped_checkpoint_known_answers(name, fp_dim, img_dim, kat):
    """best_nn_model_maccs.pth / best_nn_model.pth through the oracle net == reference outputs."""
    g = load_golden(name)
    model = nets.build("mlp", fp_dim, img_dim)
    model.load_state_dict(_state_from_golden(g), strict=True)
    model.eval()
    for batch in (1, 4, 37, 256):
        fp, img, _ = seeded_inputs(1234 + batch, batch, fp_dim, img_dim)
        assert abs(float(fp.double().sum() + img.double().sum()) - float(g[f"input_checksum_b{batch}"])) < 1e-6
        with torch.no_grad():
            out = model(fp, img).numpy()
        np.testing.assert_allclose(out, g[f"out_b{batch}"], rtol=0, atol=1e-6)
    if kat is not None:
        # SURVEY section 4 probe: Generator(1234), randn(4,64) then randn(4,128)
        gen = torch.Generator().manual_seed(1234)
        fp, img = torch.randn(4, fp_dim, generator=gen), torch.randn(4, img_dim, generator=gen)
        with torch.no_grad():
            np.testing.assert_allclose(model(fp, img).numpy().ravel(), kat, atol=2e-6)


def test_pca_projection_known_answer():
    g = load_golden("maccs_pca")
    y = preprocess.pca_transform(g["x"], g["mean"], g["components"])
    np.testing.assert_allclose(y, g["y"], atol=2e-5)
    assert g["components"].shape == (30, 167)


def test_stacking_consumer_contract():
    with open(os.path.join(GOLDEN, "stackers.json")) as fh:
        st = json.load(fh)
    assert len(st) == 5
    for v in st.values():
        assert len(v["coef"]) == 3          # NN column is feature 0 of a 3-feature linear stacker
    np.testing.assert_allclose(st["stacked_model_maccs_opt.pkl"]["coef"], [0.1407, 0.9479, 0.0776], atol=1e-4)


CASES = [
    ("tcnn_maccs", "tcnn", 167, (1, 2, 5, 32, 67), 32, 2),
    ("tcnn_nofusion_maccs", "tcnn_nofusion", 167, (4,), 8, 1),
    ("tcnn_big_maccs", "tcnn_big", 167, (3,), 4, 1),
    ("mlp_more", "mlp_more", 64, (1, 33), 16, 2),
    ("mlp_rdkit", "mlp_rdkit", 64, (9,), 16, 2),
    ("mlp_opt", "mlp", 64, (9,), 16, 2),
    ("tcnn_morgan", "tcnn", 2048, (3, 32), 8, 1),
]


@pytest.mark.parametrize("name,variant,fp_dim,batches,train_batch,steps", CASES)
def test_oracle_net_reproduces_reference_golden(name, variant, fp_dim, batches, train_batch, steps):
    """Same seed => same init => same outputs, losses, gradient norms and post-step parameters as the
    reference class instance that generated the fixture."""
    g = load_golden(name)
    assert str(g["torch_version"]) == torch.__version__, "fixtures were generated with another torch build"
    img_side = int(g["img_side"])
    img_dim = IMG if variant.startswith("tcnn") else img_side
    torch.manual_seed(int(g["init_seed"]))
    model = nets.build(variant, fp_dim, img_side)
    model.eval()
    for b in batches:
        fp, img, _ = seeded_inputs(100 + b, b, fp_dim, img_dim)
        assert abs(float(fp.double().sum() + img.double().sum()) - float(g[f"input_checksum_b{b}"])) < 1e-6
        with torch.no_grad():
            np.testing.assert_allclose(model(fp, img).numpy(), g[f"out_b{b}"], rtol=0, atol=1e-5)
    nets.zero_dropout(model)
    model.train()
    opt = torch.optim.AdamW(model.parameters(), lr=1e-4, weight_decay=1e-5)
    losses = []
    for step in range(steps):
        fp, img, y = seeded_inputs(500 + step, train_batch, fp_dim, img_dim)
        opt.zero_grad()
        loss = torch.nn.functional.mse_loss(model(fp, img).squeeze(), y)
        loss.backward()
        if step == 0:
            for k, p in model.named_parameters():
                ref = float(g["gradnorm:" + k])
                assert abs(float(p.grad.double().norm()) - ref) <= 1e-4 * max(ref, 1e-3), k
        opt.step()
        losses.append(float(loss))
    np.testing.assert_allclose(losses, g["losses"], rtol=1e-5)
    for k, p in model.state_dict().items():
        ref = float(g["after:" + k])
        assert abs(float(p.double().sum()) - ref) <= 1e-4 * max(abs(ref), 1.0), k


@pytest.mark.skipif(not rc.available(), reason="/root/reference is only present in the build container")
@pytest.mark.parametrize("variant,fp_dim,img_side", [
    ("tcnn", 167, 128), ("tcnn_first", 167, 128), ("tcnn_20250108", 167, 128), ("tcnn_nofusion", 167, 128),
    ("tcnn_big", 167, 128), ("mlp", 64, 128), ("mlp_morgan", 128, 256), ("mlp_rdkit", 64, 128), ("mlp_more", 64, 128),
])
def test_oracle_equals_reference_classes(variant, fp_dim, img_side):
    """The reference's own class (AST-loaded, unmodified) and the oracle restatement: identical
    state_dict keys/shapes, identical seeded init, identical eval outputs."""
    torch.manual_seed(3)
    ref = rc.load(variant).MixedInputModel(fp_dim, img_side)
    torch.manual_seed(3)
    ours = nets.build(variant, fp_dim, img_side)
    ref_sd, our_sd = ref.state_dict(), ours.state_dict()
    assert list(ref_sd) == list(our_sd)
    for k in ref_sd:
        assert ref_sd[k].shape == our_sd[k].shape and torch.equal(ref_sd[k], our_sd[k]), k
    img_dim = IMG if variant.startswith("tcnn") else img_side
    fp, img, _ = seeded_inputs(11, 5, fp_dim, img_dim)
    ref.eval(), ours.eval()
    with torch.no_grad():
        np.testing.assert_allclose(ours(fp, img).numpy(), ref(fp, img).numpy(), rtol=0, atol=1e-6)


def test_input_contract_oracle():
    rng = np.random.default_rng(0)
    bits = (rng.random((7, 167)) < 0.25).astype(np.uint8)
    bits[:, 0] = 0
    bits[3] = 0                                   # constant row: sklearn maps std 0 -> 1
    packed = preprocess.pack_bits(bits)
    assert packed.shape == (7, 21)
    np.testing.assert_array_equal(preprocess.unpack_bits(packed, 167), bits)
    z = preprocess.unpack_zscore(packed, 167)
    assert z.dtype == np.float32 and np.all(z[3] == 0)
    np.testing.assert_allclose(z[0].mean(), 0, atol=1e-6)
    np.testing.assert_allclose(z[0].std(), 1, atol=1e-5)
    try:
        from sklearn.preprocessing import StandardScaler
    except Exception:
        return
    for r in (0, 3, 5):   # the reference's exact call: Descriptors/multi_input_data_preprocess_maccs_opt.py:121-124
        ref = StandardScaler().fit_transform(bits[r].astype(np.float64).reshape(-1, 1)).ravel().astype(np.float32)
        np.testing.assert_array_equal(z[r], ref)


def test_chunked_standardisation_oracle_equals_sklearn():
    """N3: oracle/preprocess.standardize_chunks restates ``scaler.fit_transform`` per block of 100 molecules
    (Descriptors/multi_input_data_preprocess_maccs_opt_IsolationForest_fixed_1.py:86-101); checked here against sklearn's
    StandardScaler itself, driven exactly as the reference drives it (one scaler object, fit_transform per block, MACCS and
    pixel columns side by side), including constant columns (MACCS bit 0, white border pixels) and a ragged last block."""
    from sklearn.preprocessing import StandardScaler
    rng = np.random.default_rng(3)
    n = 257
    maccs = (rng.random((n, 167)) < 0.25).astype(np.uint8)
    maccs[:, 0] = 0
    pixels = rng.random((n, 600)).astype(np.float32)
    pixels[:, :40] = 1.0                                        # constant (white) pixels
    pixels[:, 40:60] = np.float32(1.0) - (rng.random((n, 20)) < 0.01) * np.float32(0.37)   # almost constant
    scaler = StandardScaler()
    want = []
    for i in range(0, n, 100):
        block = np.hstack([maccs[i:i + 100], pixels[i:i + 100]])
        want.extend(scaler.fit_transform(block))
    want = np.array(want, dtype=np.float32)
    got = preprocess.standardize_chunks(np.hstack([maccs, pixels]).astype(np.float32), 100)
    np.testing.assert_array_equal(got, want)
