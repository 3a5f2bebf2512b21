import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
ORACLE = os.path.join(ROOT, "oracle")
if ORACLE not in sys.path:
    sys.path.insert(0, ORACLE)
GOLDEN = os.path.join(ROOT, "tests", "golden")

# The oracle is the checker: on a GPU box make its contractions independent of how that host runs fp32 GEMMs
# (oracle/zuko/nn.py; hosts of this pool were seen to differ by 4e-4 on the same seeded forward).  Without a GPU
# (the build container) the oracle keeps zuko's plain fp32 arithmetic, the mode its goldens were recorded in.
import torch  # noqa: E402
import zuko.nn as _oracle_nn  # noqa: E402  (oracle/zuko, the restatement -- not the third-party package)
_oracle_nn.MATMUL_FP64 = bool(torch.cuda.is_available())


@pytest.fixture
def faithful_fp32_oracle():
    """Plain fp32 oracle arithmetic for tests that compare two fp32 CPU computations with each other."""
    old = _oracle_nn.MATMUL_FP64
    _oracle_nn.MATMUL_FP64 = False
    yield
    _oracle_nn.MATMUL_FP64 = old


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_collection_modifyitems(config, items):
    """GPU tests fail loudly (not skip) on a box whose GPU/extension is broken, but are skipped
    when there is simply no CUDA device and the run did not ask for them with -m gpu."""
    import torch
    if torch.cuda.is_available():
        return
    selected = config.getoption("-m") or ""
    if "gpu" in selected and "not gpu" not in selected:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden():
    def load(name):
        return dict(np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False))
    return load


def flow_param_list(g, prefix):
    """Golden flow parameters (module order) stored as <prefix>p000, p001, ..."""
    keys = sorted(k for k in g if k.startswith(prefix + "p") and k[len(prefix) + 1:].isdigit())
    return [g[k] for k in keys]
