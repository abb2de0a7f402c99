"""Shared fixtures.  ``-m "not gpu"`` runs everywhere (oracle vs golden vectors, host logic, C-ABI symbol checks,
gloo world-size-2 sharding); ``-m gpu`` are the parity tests proper and need a B200."""
import os
import sys

import numpy as np
import pandas as pd
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")
REFERENCE = "/root/reference"          # only present in the build container; tests that import it skip elsewhere


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def test_counts():
    """examples/test.csv of the reference (500 cells x 3000 genes), shipped as tests/golden/test_counts.npz."""
    z = np.load(os.path.join(GOLDEN, "test_counts.npz"))
    return pd.DataFrame(z["counts"].astype(np.float64), index=z["cells"].astype(object),
                        columns=z["genes"].astype(object))


@pytest.fixture(scope="session")
def golden_partition():
    def load(name):
        z = np.load(os.path.join(GOLDEN, "partition_{}.npz".format(name)))
        pred = np.split(z["pred_flat"], np.cumsum(z["pred_len"])[:-1])
        return dict(targets=z["targets"], predictors=pred, test_rows=z["test_rows"], train_rows=z["train_rows"])
    return load


def synthetic_counts(n_cells, n_genes, seed=0, rank=8):
    """Low-rank overdispersed counts (SURVEY.md 8d generator, small)."""
    rng = np.random.default_rng(seed)
    Z = rng.gamma(2.0, 0.5, size=(n_cells, rank))
    W = rng.gamma(0.3, 1.0, size=(n_genes, rank))
    scale = rng.lognormal(0.0, 1.0, size=n_genes)
    lam = (Z @ W.T) * scale * (4.0 / rank)
    raw = rng.poisson(lam).astype(np.float64)
    raw[0, 0] = max(raw[0, 0], 12.0)
    return pd.DataFrame(raw, index=["c{}".format(i) for i in range(n_cells)],
                        columns=["g{}".format(j) for j in range(n_genes)])
