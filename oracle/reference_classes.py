"""Load the reference's model classes straight from its scripts (test infrastructure).

The reference scripts cannot be imported: they ``pickle.load`` an un-shipped file and
train at module level.  We parse the script, keep only its ``ClassDef`` nodes and exec
those with ``torch`` / ``nn`` / ``Dataset`` in scope, so the classes that run are the
reference's own source, unmodified, without copying it into this repo.

Only usable where ``/root/reference`` exists (the build container).  The GPU box has no
such tree: tests that need this module skip there and rely on ``tests/golden``.
"""
from __future__ import annotations

import ast
import os
from types import SimpleNamespace

REFERENCE_ROOT = os.environ.get("BBBP_REFERENCE_ROOT", "/root/reference")

# variant name -> script that declares it (paths relative to REFERENCE_ROOT)
SCRIPTS = {
    "tcnn": "Models/multi_input_data_regression_opt_transformer_cnn_20250113.py",
    "tcnn_first": "Models/multi_input_data_regression_opt_transformer_cnn.py",
    "tcnn_20250108": "Models/multi_input_data_regression_opt_transformer_cnn_20250108.py",
    "tcnn_big": "Models/multi_input_data_regression_opt_transformer_cnn_opt_20250107_network.py",
    "tcnn_nofusion": "Descriptors/multi_input_data_regression_opt_round_2_transformer_cnn.py",
    "mlp": "Models/multi_input_data_regression_opt_transformer_cnn_opt.py",
    "mlp_morgan": "Models/multi_input_data_regression_opt_transformer_cnn_morgan.py",
    "mlp_rdkit": "Models/multi_input_data_regression_opt_transformer_cnn_rdkit.py",
    "mlp_more": "Models/multi_input_data_regression_opt_transformer_cnn_opt_more.py",
}


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "Models"))


def load(variant: str) -> SimpleNamespace:
    """Return a namespace holding every class the variant's script declares."""
    import torch
    import torch.nn as nn
    from torch.utils.data import Dataset

    path = os.path.join(REFERENCE_ROOT, SCRIPTS[variant])
    with open(path, "r", encoding="utf-8") as fh:
        tree = ast.parse(fh.read(), filename=path)
    keep = [node for node in tree.body if isinstance(node, ast.ClassDef)]
    module = ast.Module(body=keep, type_ignores=[])
    scope = {"torch": torch, "nn": nn, "Dataset": Dataset, "__name__": "reference_" + variant}
    exec(compile(module, path, "exec"), scope)
    return SimpleNamespace(**{k: v for k, v in scope.items() if isinstance(v, type)})


def artefact(relpath: str) -> str:
    return os.path.join(REFERENCE_ROOT, relpath)
