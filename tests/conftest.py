import os
import sys
import warnings

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

warnings.filterwarnings("ignore", message=".*enable_nested_tensor.*")

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    # the C-ABI library must exist before anything imports bbbp_b200 (it is git-ignored, so a fresh
    # checkout has none); building is not using: no kernel runs here
    import importlib.util
    spec = importlib.util.spec_from_file_location(
        "_bbbp_build", os.path.join(ROOT, "bbbp-multi-modal-deep-ensemble-framework_b200", "build.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    mod.build_library()      # no-op unless the library is missing or older than a source / header (build.is_stale)


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


def load_golden(name):
    import numpy as np
    return np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False)


def seeded_inputs(seed, batch, fp_dim, img_dim):
    """Same recipe as oracle/make_golden.py:seeded_inputs."""
    import torch
    g = torch.Generator().manual_seed(seed)
    fp = torch.randn(batch, fp_dim, generator=g)
    img = torch.randn(batch, img_dim, generator=g)
    y = torch.randn(batch, generator=g) * 0.75 - 0.1
    return fp, img, y


@pytest.fixture(scope="session")
def cuda_device():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import bbbp_b200
    bbbp_b200.ops.require_device()
    return torch.device("cuda:0")
