"""GPU: size-independent properties at the sizes and shapes of BASELINE.json's configs (the oracle needs
minutes per MCMC step there, so parity against it is checked at small sizes elsewhere):

  * flow round trip / log-det antisymmetry on every sweep path the shape selects (reference
    tests/test_flow.py:75-88,153-166 at 1e-5 in fp32 -- scaled here by the magnitude of the values);
  * one `_mutate`-sized call of the reference-facing kernel (pocomc/mcmc.py:8-183) must return a
    SELF-CONSISTENT state: x == scaler.inverse(u), logdetj its log-determinant, logl == loglike(x),
    logp == logprior(x) for every particle, accepted or not; the acceptance rate must be a probability
    and the adapted proposal scale must respect its cap (mcmc.py:152);
  * the likelihood is only ever called on finite rows and the call counter adds up (mcmc.py:106,115-121).
"""
import numpy as np
import pytest
import torch
from scipy.stats import norm, uniform

pytestmark = pytest.mark.gpu

CONFIGS = {
    # name: (n_dim, n_particles, flow, likelihood factory, prior factor, bounded)
    "cfg2_gauss32": (32, 10_000, "maf6", "gauss", norm(0.0, 3.0)),
    "cfg3_mixture50": (50, 50_000, "maf6", "mixture", uniform(-10.0, 20.0)),
    "cfg4_rosen100": (100, 20_000, "maf3", "rosen", uniform(-10.0, 20.0)),     # 1/10 of one rank's shard, big-H sweep kernel
    "cfg5_funnel200": (200, 4_096, "maf3", "funnel", uniform(-30.0, 60.0)),    # H = 1024, Student-t proposal nu = 5
}


def _likelihood(kind, d):
    from pocomc_b200 import synthetic as S
    return dict(gauss=lambda: S.CorrelatedGaussian(d), mixture=lambda: S.GaussianMixture(), rosen=lambda: S.Rosenbrock(),
                funnel=lambda: S.Funnel())[kind]()


@pytest.mark.parametrize("name", list(CONFIGS))
def test_flow_round_trip_at_config_shapes(name):
    import pocomc_b200 as pc
    d, n, preset, _, _ = CONFIGS[name]
    torch.manual_seed(d)
    f = pc.Flow(d, preset)
    with torch.no_grad():
        f.flow.raw.mul_(1.3)                      # away from the near-identity initialisation
    x = torch.randn(min(n, 20_000), d)
    with torch.no_grad():
        z, l_fwd = f.forward(x)
        xb, l_inv = f.inverse(z)
    scale = max(1.0, float(z.abs().max()))
    assert torch.isfinite(z).all() and torch.isfinite(l_fwd).all()
    np.testing.assert_allclose(xb.numpy(), x.numpy(), atol=2e-5 * scale, rtol=2e-5)
    np.testing.assert_allclose(l_inv.numpy(), -l_fwd.numpy(), atol=2e-5 * max(1.0, float(l_fwd.abs().max())), rtol=2e-5)


@pytest.mark.parametrize("name", list(CONFIGS))
def test_mutate_returns_self_consistent_state(name):
    import pocomc_b200 as pc
    from pocomc_b200 import config, mcmc
    d, n, preset, like_kind, factor = CONFIGS[name]
    rng = np.random.default_rng(d)
    like = _likelihood(like_kind, d)
    prior = pc.Prior([factor] * d)
    calls = []

    def loglike(x):
        assert np.isfinite(x).all()
        calls.append(len(x))
        return like(x), None

    x0 = prior.rvs(n) if like_kind != "gauss" else rng.normal(size=(n, d))
    if like_kind == "mixture":
        x0 = rng.normal(size=(n, d)) * 0.5 + 1.5 * rng.choice([-1.0, 1.0], size=(n, 1))
    if like_kind == "funnel":
        x0 = rng.normal(size=(n, d))
    scaler = pc.scaler.Reparameterize(d, bounds=prior.bounds)
    scaler.fit(prior.rvs(4096))
    u0 = scaler.forward(x0)
    ldj0 = scaler.inverse(u0)[1]
    torch.manual_seed(1)
    flow = pc.Flow(d, preset)
    theta0 = pc.tools.flow_numpy_wrapper(flow).forward(u0)[0]
    geo = pc.geometry.Geometry()
    geo.fit(theta0[:4096].astype(np.float64))
    if like_kind == "funnel":
        geo.t_nu = 5.0
    config.set_rng_mode("device")
    try:
        state = dict(u=u0, x=x0, logdetj=ldj0, logl=like(x0), logp=prior.logpdf(x0), beta=0.5, blobs=None)
        res = mcmc.preconditioned_pcn(dict(state), dict(loglike=loglike, logprior=prior.logpdf, scaler=scaler, flow=flow,
                                                        theta_geometry=geo, u_geometry=geo),
                                      dict(n_max=3, n_steps=10 ** 6, progress_bar=None, proposal_scale=2.38 / d ** 0.5, seed=5))
    finally:
        config.set_rng_mode("host")
    assert res["steps"] == 3 and 0.0 <= res["accept"] <= 1.0
    assert 0.0 < res["proposal_scale"] <= min(2.38 / d ** 0.5, 0.99) + 1e-15
    assert res["calls"] == sum(calls) <= 3 * n
    x_chk, ldj_chk = scaler.inverse(res["u"])
    np.testing.assert_allclose(res["x"], x_chk, rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(res["logdetj"], ldj_chk, rtol=1e-12, atol=1e-10)
    np.testing.assert_allclose(res["logl"], like(res["x"]), rtol=1e-12, atol=1e-9)
    np.testing.assert_allclose(res["logp"], prior.logpdf(res["x"]), rtol=1e-12, atol=1e-9)
    moved = np.any(res["x"] != x0, axis=1)
    assert moved.any() or res["accept"] < 1e-3           # accept = mean Metropolis probability of the last step
