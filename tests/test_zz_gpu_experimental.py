"""GPU parity of the EXPERIMENTAL sweep kernels (opt-in variants, never the default path).  Kept in a file that sorts
last so that a problem here cannot stop the parity tests of the default path under ``pytest -x``.
csrc/flow_tip.cu: bulk/tip sweep (validated on B200) and its register-blocked variants (not yet run: behind
PMC_B200_EXPERIMENTAL=1)."""
import numpy as np
import pytest
import torch

import flow_ref as F
from test_gpu_flow import _mine

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("preset,d,n", [("maf6", 32, 1000), ("maf3", 10, 77), ("maf3", 21, 300), ("maf6", 50, 64), ("maf3", 6, 9)])
def test_bulk_tip_sweep_matches_oracle(preset, d, n):
    """config.sweep_variant = "tip": same parity bar as the default sweep kernel (5e-5), both directions."""
    from pocomc_b200 import config
    torch.manual_seed(d * 3 + n)
    ref = F.make_flow(d, preset)
    old = config.sweep_variant, config.forward_path
    config.sweep_variant, config.forward_path = "tip", "sweep"
    try:
        f = _mine(preset, d, [p.detach().numpy() for p in ref.parameters()])
        assert int(f.flow._meta_host[22]) == 5                     # made_layout.M_VERSION: the bulk/tip stream
        x = torch.randn(n, d)
        with torch.no_grad():
            z_ref, l_ref = ref().transform.call_and_ladj(x)
            xi_ref, li_ref = ref().transform.inv.call_and_ladj(z_ref)
            z, l = f.forward(x)
            xi, li = f.inverse(z_ref)
    finally:
        config.sweep_variant, config.forward_path = old
    tol = dict(rtol=5e-5, atol=5e-5)
    np.testing.assert_allclose(z.numpy(), z_ref.numpy(), **tol)
    np.testing.assert_allclose(l.numpy(), l_ref.numpy(), **tol)
    np.testing.assert_allclose(xi.numpy(), xi_ref.numpy(), **tol)
    np.testing.assert_allclose(li.numpy(), li_ref.numpy(), **tol)


@pytest.mark.skipif(__import__("os").environ.get("PMC_B200_EXPERIMENTAL") != "1",
                    reason="register-blocked bulk/tip variants (PMC_TIP_PPL = 2 | 4) have not been run on a GPU yet: "
                           "set PMC_B200_EXPERIMENTAL=1")
@pytest.mark.parametrize("ppl", [2, 4])
@pytest.mark.parametrize("preset,d,n", [("maf6", 32, 1000), ("maf3", 10, 77), ("maf3", 21, 300), ("maf3", 6, 9)])
def test_bulk_tip_sweep_register_blocked_variants(preset, d, n, ppl, monkeypatch):
    from pocomc_b200 import config
    monkeypatch.setenv("PMC_TIP_PPL", str(ppl))
    torch.manual_seed(d * 3 + n)
    ref = F.make_flow(d, preset)
    old = config.sweep_variant, config.forward_path
    config.sweep_variant, config.forward_path = "tip", "sweep"
    try:
        f = _mine(preset, d, [p.detach().numpy() for p in ref.parameters()])
        x = torch.randn(n, d)
        with torch.no_grad():
            z_ref, l_ref = ref().transform.call_and_ladj(x)
            xi_ref, li_ref = ref().transform.inv.call_and_ladj(z_ref)
            z, l = f.forward(x)
            xi, li = f.inverse(z_ref)
    finally:
        config.sweep_variant, config.forward_path = old
    tol = dict(rtol=5e-5, atol=5e-5)
    np.testing.assert_allclose(z.numpy(), z_ref.numpy(), **tol)
    np.testing.assert_allclose(l.numpy(), l_ref.numpy(), **tol)
    np.testing.assert_allclose(xi.numpy(), xi_ref.numpy(), **tol)
    np.testing.assert_allclose(li.numpy(), li_ref.numpy(), **tol)


@pytest.mark.skipif(__import__("os").environ.get("PMC_B200_EXPERIMENTAL") != "1",
                    reason="split-K hidden GEMMs of the training kernel (PMC_TRAIN_SPLITK=1) have not been run on a GPU "
                           "yet: set PMC_B200_EXPERIMENTAL=1")
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
