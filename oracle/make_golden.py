"""Generate tests/golden/*.npz from the REFERENCE's own classes and artefacts.

Run in the build container only (needs /root/reference):   python -m oracle.make_golden
TEST INFRASTRUCTURE (see oracle/__init__.py).  The fixtures pin oracle/nets.py and the
CUDA product on boxes where /root/reference does not exist.

Every fixture records ``torch.__version__``.  Inputs are regenerated from CPU generator
seeds; an ``input_checksum`` guards against RNG drift between torch builds.
"""
from __future__ import annotations

import json
import os
import pickle
import warnings

import numpy as np
import torch

from . import nets, reference_classes as rc

GOLDEN = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
warnings.filterwarnings("ignore", message="enable_nested_tensor")


def seeded_inputs(seed: int, batch: int, fp_dim: int, img_dim: int):
    """The input recipe shared by the generator and the tests."""
    g = torch.Generator().manual_seed(seed)
    fp = torch.randn(batch, fp_dim, generator=g)
    img = torch.randn(batch, img_dim, generator=g)
    y = torch.randn(batch, generator=g) * 0.75 - 0.1
    return fp, img, y


def _np(t):
    return t.detach().cpu().numpy()


def _load_pickle(path):
    """The shipped .pkl files are a mix of plain pickles and joblib dumps."""
    try:
        with open(path, "rb") as fh:
            return pickle.load(fh)
    except pickle.UnpicklingError:
        import joblib
        return joblib.load(path)


def checkpoint_kat(name: str, pth: str, fp_dim: int, img_dim: int):
    """Known-answer vectors from a shipped MLP-family checkpoint."""
    model = rc.load("mlp").MixedInputModel(fp_dim, img_dim)
    state = torch.load(rc.artefact(pth), map_location="cpu")
    model.load_state_dict(state, strict=True)
    model.eval()
    out = {"torch_version": torch.__version__, "fp_dim": fp_dim, "img_dim": img_dim}
    for k, v in state.items():
        out["param:" + k] = _np(v)
    for batch in (1, 4, 37, 256):
        fp, img, _ = seeded_inputs(1234 + batch, batch, fp_dim, img_dim)
        with torch.no_grad():
            out[f"out_b{batch}"] = _np(model(fp, img))
        out[f"input_checksum_b{batch}"] = np.float64(fp.double().sum() + img.double().sum())
    np.savez_compressed(os.path.join(GOLDEN, name + ".npz"), **out)
    print(name, "b4 ->", out["out_b4"].ravel())


def pca_kat():
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        pca = _load_pickle(rc.artefact("Models/maccs_pca.pkl"))
    g = torch.Generator().manual_seed(77)
    x = torch.randn(19, pca.components_.shape[1], generator=g).numpy()
    # The pickle predates sklearn 1.9 (its transform() trips on a missing attribute), so apply
    # sklearn's published PCA.transform formula to the pickled attributes in float64.
    assert not pca.whiten
    y = (x.astype(np.float64) - pca.mean_.astype(np.float64)) @ pca.components_.astype(np.float64).T
    np.savez_compressed(os.path.join(GOLDEN, "maccs_pca.npz"), components=pca.components_.astype(np.float32),
                        mean=pca.mean_.astype(np.float32), x=x, y=y.astype(np.float32))
    print("maccs_pca", pca.components_.shape)


def stacking_contract():
    """The consumer of the NN column: 3-coefficient linear stackers (NN is feature 0)."""
    out = {}
    for f in ("stacked_model.pkl", "stacked_model_maccs_opt.pkl", "stacked_model_maccs_multiattention.pkl",
              "stacked_model_morgan.pkl", "stacked_model_rdkit.pkl"):
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            m = _load_pickle(rc.artefact("Models/" + f))
        out[f] = {"type": type(m).__name__, "coef": np.asarray(m.coef_).ravel().tolist(), "intercept": float(np.ravel(m.intercept_)[0])}
    with open(os.path.join(GOLDEN, "stackers.json"), "w") as fh:
        json.dump(out, fh, indent=1)
    print("stackers", {k: v["coef"] for k, v in out.items()})


def seeded_net_kat(name: str, variant: str, fp_dim: int, img_side: int, batches, train_batch: int, init_seed: int = 0,
                   steps: int = 2):
    """Random-init reference instance: eval forward at several batch sizes, then a
    dropout-free train-mode step sequence (loss, gradient norms, parameter checksums)."""
    img_dim = 3 * 128 * 128 if variant.startswith("tcnn") else img_side
    torch.manual_seed(init_seed)
    model = rc.load(variant).MixedInputModel(fp_dim, img_side)
    out = {"torch_version": torch.__version__, "variant": variant, "fp_dim": fp_dim, "img_side": img_side,
           "init_seed": init_seed, "train_batch": train_batch}
    model.eval()
    for batch in batches:
        fp, img, _ = seeded_inputs(100 + batch, batch, fp_dim, img_dim)
        with torch.no_grad():
            out[f"out_b{batch}"] = _np(model(fp, img))
        out[f"input_checksum_b{batch}"] = np.float64(fp.double().sum() + img.double().sum())
    # training trajectory with every dropout off and BatchNorm in batch-stat mode
    nets.zero_dropout(model)
    model.train()
    opt = torch.optim.AdamW(model.parameters(), lr=1e-4, weight_decay=1e-5)   # 20250113.py:172
    losses = []
    for step in range(steps):
        fp, img, y = seeded_inputs(500 + step, train_batch, fp_dim, img_dim)
        opt.zero_grad()
        loss = torch.nn.functional.mse_loss(model(fp, img).squeeze(), y)
        loss.backward()
        if step == 0:
            for k, p in model.named_parameters():
                out["gradnorm:" + k] = np.float64(p.grad.double().norm())
        opt.step()
        losses.append(float(loss))
    out["losses"] = np.asarray(losses, dtype=np.float64)
    for k, p in model.state_dict().items():
        out["after:" + k] = np.float64(p.double().sum())
    model.eval()
    fp, img, _ = seeded_inputs(900, 7, fp_dim, img_dim)
    with torch.no_grad():
        out["out_after_b7"] = _np(model(fp, img))
    np.savez_compressed(os.path.join(GOLDEN, name + ".npz"), **out)
    print(name, "losses", losses, "out_b%d[0]" % batches[0], out[f"out_b{batches[0]}"].ravel()[:3])


def main():
    os.makedirs(GOLDEN, exist_ok=True)
    assert rc.available(), "needs /root/reference"
    torch.set_num_threads(8)
    checkpoint_kat("mlp_ckpt_maccs", "Models/best_nn_model_maccs.pth", 64, 128)
    checkpoint_kat("mlp_ckpt_morgan", "Models/best_nn_model.pth", 128, 256)
    pca_kat()
    stacking_contract()
    seeded_net_kat("tcnn_maccs", "tcnn", 167, 128, batches=(1, 2, 5, 32, 67), train_batch=32)
    seeded_net_kat("tcnn_morgan", "tcnn", 2048, 128, batches=(3, 32), train_batch=8, steps=1)
    seeded_net_kat("tcnn_nofusion_maccs", "tcnn_nofusion", 167, 128, batches=(4,), train_batch=8, steps=1)
    seeded_net_kat("tcnn_big_maccs", "tcnn_big", 167, 128, batches=(3,), train_batch=4, steps=1)
    seeded_net_kat("mlp_more", "mlp_more", 64, 128, batches=(1, 33), train_batch=16)
    seeded_net_kat("mlp_rdkit", "mlp_rdkit", 64, 128, batches=(9,), train_batch=16)
    seeded_net_kat("mlp_opt", "mlp", 64, 128, batches=(9,), train_batch=16)


if __name__ == "__main__":
    main()
