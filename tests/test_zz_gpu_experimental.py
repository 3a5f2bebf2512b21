"""GPU parity of opt-in kernel variants (never the default path).  Kept in a file that sorts last so that a problem
here cannot stop the parity tests of the default path under ``pytest -x``."""
import numpy as np
import pytest
import torch

import flow_ref as F
from test_gpu_flow import _mine

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("preset,d,n,weighted", [("maf6", 32, 300, True), ("maf3", 42, 64, True), ("maf3", 21, 33, False),
                                                 ("maf6", 50, 512, True)])
def test_training_kernel_split_k_variant(preset, d, n, weighted, monkeypatch):
    """csrc/flow_train.cu with PMC_TRAIN_SPLITK=1: same loss / gradient bars as the default tiling."""
    from pocomc_b200.flow import _FitEngine
    monkeypatch.setenv("PMC_TRAIN_SPLITK", "1")
    torch.manual_seed(d + n)
    ref = F.make_flow(d, preset)
    f = _mine(preset, d, [p.detach().numpy() for p in ref.parameters()])
    x = torch.randn(n + 50, d) * 1.2 + 0.1
    w = torch.rand(n + 50) + 0.05
    rows = torch.randperm(n + 50)[:n]
    lp = ref().log_prob(x[rows])
    loss_ref = (-lp * w[rows] * 1000.0).sum() / w[rows].sum() if weighted else -lp.sum()
    loss_ref.backward()
    gref = torch.cat([p.grad.reshape(-1) for p in ref.parameters()]).numpy()
    eng = _FitEngine(f.flow)
    assert eng.fused
    eng.load(x.cuda(), w.cuda())
    loss, g = eng.loss_and_grad(rows, weighted)
    np.testing.assert_allclose(loss, float(loss_ref.detach()), rtol=1e-5)
    g = g.cpu().numpy()
    np.testing.assert_allclose(g, gref, rtol=2e-4, atol=1e-4 * np.abs(gref).max())
    assert np.all(g[gref == 0] == 0)
