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
