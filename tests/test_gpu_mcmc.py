"""GPU parity: the four MCMC kernels through the reference's dict protocol against runs of the
reference's own pocomc/mcmc.py (tests/golden/mcmc_*.npz), fed the same legacy np.random stream.

Tolerances: the reference state is f64 but every value that went through the fp32 flow carries
~1e-6 relative rounding differences (different fp32 summation order than MKL), so u/x/logdetj are
compared at 2e-5; accept decisions must agree exactly unless |r - alpha| is within that noise.
"""
import numpy as np
import pytest
import torch

import flow_ref as F
import smc_ref as O
from conftest import flow_param_list

pytestmark = pytest.mark.gpu


class _Geo:
    pass


def _setup(g, preset="maf3"):
    from pocomc_b200.flow import Flow
    from pocomc_b200.scaler import Reparameterize
    from scipy.stats import norm, uniform
    d = g["x"].shape[1]
    flow = Flow(d, preset)
    flat = np.concatenate([p.reshape(-1) for p in flow_param_list(g, "")])
    with torch.no_grad():
        flow.flow.raw.copy_(torch.from_numpy(flat).to(flow.flow.raw.device))
    scaler = Reparameterize(d, bounds=np.stack([g["low"], g["high"]], 1))
    scaler.mu, scaler.sigma = g["mu"], g["sigma"]
    Ci = g["Cinv"]
    dists = [uniform(-6, 12) if k else norm(0, 3) for k in g["prior_kind"]]

    def loglike(x):
        return -0.5 * np.einsum("ki,ij,kj->k", x, Ci, x), None

    def logprior(x):
        return sum(dd.logpdf(x[:, i]) for i, dd in enumerate(dists))

    return d, flow, scaler, loglike, logprior


_KEYS = ["tpcn_flow_nufit", "tpcn_flow_nu5", "rwm_flow_nufit", "tpcn_nufit", "tpcn_nu5", "rwm_nufit"]


@pytest.mark.parametrize("tag,key", [(t, k) for t in ("free", "bounded") for k in _KEYS] + [("nsf", k) for k in _KEYS[:3]])
def test_kernels_match_reference_runs(golden, tag, key):
    """tag "nsf": the reference's default flow family (neural spline flow, sampler.py:169) through the flow-preconditioned
    kernels; the spline's fp32 noise is ~10x the affine map's (tests/test_gpu_flow.py), so its bar is 2e-4."""
    from pocomc_b200 import mcmc, config
    g = golden("mcmc_" + tag)
    preset = "nsf3" if tag == "nsf" else "maf3"
    d, flow, scaler, loglike, logprior = _setup(g, preset)
    config.set_rng_mode("host")
    kind = key.rsplit("_", 1)[0]
    fn = dict(tpcn_flow=mcmc.preconditioned_pcn, rwm_flow=mcmc.preconditioned_rwm, tpcn=mcmc.pcn, rwm=mcmc.rwm)[kind]
    geo = _Geo()
    geo.t_mean, geo.t_cov, geo.t_nu = g[f"{key}_t_mean"], g[f"{key}_t_cov"], float(g[f"{key}_t_nu"])
    geo.normal_cov = g[f"{key}_normal_cov"]
    state = dict(u=g["u"], x=g["x"], logdetj=g["logdetj"], logl=g["logl"], logp=g["logp"], beta=float(g["beta"]), blobs=None)
    fd = dict(loglike=loglike, logprior=logprior, scaler=scaler, flow=flow, u_geometry=geo, theta_geometry=geo)
    od = dict(n_max=6, n_steps=3, progress_bar=None, proposal_scale=2.38 / d ** 0.5)
    np.random.seed(int(g[f"{key}_seed"]))
    res = fn(state, fd, od)
    assert res["steps"] == int(g[f"{key}_out_steps"])
    tol = dict(rtol=2e-5, atol=2e-5) if kind.endswith("flow") else dict(rtol=1e-10, atol=1e-10)
    if tag == "nsf":
        tol = dict(rtol=2e-4, atol=2e-4)
    # Accept decisions are discrete: a row may leave the reference trajectory ONLY where some step's uniform draw sat
    # within the fp32 flow noise of its acceptance probability (|r - alpha| < 1e-5).  Those rows are identified with
    # the oracle (replaying the recorded noise, per-step alpha in its trace); every other row must match to rounding.
    marginal = np.zeros(len(g["x"]), bool)
    if kind.endswith("flow"):
        trace = []
        steps = [O.Noise(None if g[f"{key}_g"].ndim < 2 else g[f"{key}_g"][i], g[f"{key}_z"][i], g[f"{key}_r"][i])
                 for i in range(len(g[f"{key}_r"]))]
        ref_flow = F.make_flow(d, preset)
        with torch.no_grad():
            for p_, v in zip(ref_flow.parameters(), flow_param_list(g, "")):
                p_.copy_(torch.from_numpy(v))
        sp = O.ScalerParams(low=g["low"], high=g["high"], mu=g["mu"], sigma=g["sigma"])
        O.mcmc_kernel(kind, state, lambda x: loglike(x)[0], logprior, sp,
                      dict(t_mean=geo.t_mean, t_cov=geo.t_cov, t_nu=geo.t_nu, normal_cov=geo.normal_cov), od,
                      flow=F.NumpyFlow(ref_flow), noise=O.ReplayNoise(steps), trace=trace)
        for i, tr in enumerate(trace):
            marginal |= np.abs(g[f"{key}_r"][i] - tr["alpha"]) < (1e-4 if tag == "nsf" else 1e-5)
    bad = np.zeros(len(g["x"]), bool)
    for k in ("u", "x"):
        diff = np.abs(res[k] - g[f"{key}_out_{k}"]).max(axis=1)
        bad |= diff > tol["atol"] + tol["rtol"] * np.abs(g[f"{key}_out_{k}"]).max(axis=1)
    assert not np.any(bad & ~marginal), f"{int(np.sum(bad & ~marginal))} rows diverged without a marginal accept decision"
    if not bad.any():
        for k in ("logdetj", "logl", "logp"):
            np.testing.assert_allclose(res[k], g[f"{key}_out_{k}"], **tol)
        assert res["calls"] == int(g[f"{key}_out_calls"])
        for k in ("efficiency", "accept", "proposal_scale"):
            np.testing.assert_allclose(res[k], g[f"{key}_out_{k}"], rtol=1e-6 if kind.endswith("flow") else 1e-11)
    else:       # a marginal flip changes the accepted set, not the acceptance probabilities of this call's last step much
        print(f"{key}/{tag}: {int(bad.sum())} marginal row(s) flipped")


def test_single_step_operators_vs_oracle(golden):
    """proposal + Mahalanobis, scaler pull-back and Metropolis update, one operator at a time (f64: 1e-12)."""
    from pocomc_b200 import _lib
    import ctypes as C
    g = golden("mcmc_bounded")
    key = "tpcn_nu5"
    n, d = g["x"].shape
    dev = torch.device("cuda")
    mu, cov, nu = g[f"{key}_t_mean"], g[f"{key}_t_cov"], float(g[f"{key}_t_nu"])
    inv, chol = np.linalg.inv(cov), np.linalg.cholesky(cov)
    sigma = 0.37
    gg, zz, rr = g[f"{key}_g"][0], g[f"{key}_z"][0], g[f"{key}_r"][0]
    prop_ref, m_ref = O.tpcn_propose(g["u"], mu, inv, chol, nu, sigma, gg, zz)
    mp_ref = O.mahalanobis(prop_ref - mu, inv)
    t = lambda a, dt=torch.float64: torch.as_tensor(np.ascontiguousarray(a), dtype=dt).to(dev)
    ctl = np.zeros(16 + d); ctl[0] = sigma; ctl[16:] = mu
    ctl_d = t(ctl)
    prop = torch.empty((n, d), dtype=torch.float64, device=dev)
    m_cur = torch.empty(n, dtype=torch.float64, device=dev); m_prop = torch.empty_like(m_cur)
    u_d, g_d, z_d = t(g["u"]), t(gg), t(zz)
    inv_d, chol_d = t(inv.T), t(chol.T)          # keep the device copies alive across the async launch
    _lib.call("pmc_tpcn_propose", 0, _lib.ptr(u_d), _lib.ptr(ctl_d), _lib.ptr(inv_d), _lib.ptr(chol_d), nu,
              _lib.ptr(g_d), _lib.ptr(z_d), _lib.ptr(prop), None, _lib.ptr(m_cur), _lib.ptr(m_prop), n, d)
    np.testing.assert_allclose(prop.cpu().numpy(), prop_ref, rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(m_cur.cpu().numpy(), m_ref, rtol=1e-12)
    np.testing.assert_allclose(m_prop.cpu().numpy(), mp_ref, rtol=1e-12)
    # rwm proposal
    pr = torch.empty_like(prop)
    _lib.call("pmc_rwm_propose", 0, _lib.ptr(u_d), _lib.ptr(ctl_d), _lib.ptr(chol_d), _lib.ptr(z_d), _lib.ptr(pr), None, n, d)
    np.testing.assert_allclose(pr.cpu().numpy(), O.rwm_propose(g["u"], chol, sigma, zz), rtol=1e-12, atol=1e-12)
    # Metropolis update with made-up primed scalars incl. -inf / nan cases
    rng = np.random.default_rng(0)
    logl_p = g["logl"] + rng.normal(size=n); logp_p = g["logp"] + rng.normal(size=n) * 0.1
    ldj_p = g["logdetj"] + rng.normal(size=n) * 0.1
    logl_p[:3] = -np.inf; logp_p[3:5] = -np.inf; ldj_p[5] = np.nan
    beta = 0.0 if False else 0.6
    A, B = O.t_factor(mp_ref, d, nu), O.t_factor(m_ref, d, nu)
    alpha_ref = O.mh_alpha(beta, logl_p, g["logl"], logp_p, g["logp"], ldj_p, g["logdetj"], None, None, A, B)
    acc = rr < alpha_ref
    x_p = g["x"] + 1.0
    st = {k: t(g[k]) for k in ("u", "x", "logdetj", "logl", "logp")}
    parts = torch.zeros(int(_lib.load().pmc_mh_partials_size(n, d)), dtype=torch.float64, device=dev)
    alpha = torch.empty(n, dtype=torch.float64, device=dev)
    xp_d, ldjp_d, llp_d, lpp_d, rr_d = t(x_p), t(ldj_p), t(logl_p), t(logp_p), t(rr)
    _lib.call("pmc_mh_accept_update", 2, beta, nu, None, _lib.ptr(st["u"]), _lib.ptr(st["x"]), _lib.ptr(st["logdetj"]),
              _lib.ptr(st["logl"]), _lib.ptr(st["logp"]), None, _lib.ptr(prop), _lib.ptr(prop), _lib.ptr(xp_d),
              _lib.ptr(ldjp_d), _lib.ptr(llp_d), _lib.ptr(lpp_d), None, _lib.ptr(m_cur), _lib.ptr(m_prop),
              _lib.ptr(rr_d), None, _lib.ptr(alpha), _lib.ptr(parts), n, d)
    np.testing.assert_allclose(alpha.cpu().numpy(), alpha_ref, rtol=1e-11, atol=1e-300)
    exp_x = np.where(acc[:, None], x_p, g["x"])
    np.testing.assert_array_equal(st["x"].cpu().numpy(), exp_x)
    np.testing.assert_array_equal(st["logl"].cpu().numpy(), np.where(acc, logl_p, g["logl"]))
    # controller: sigma adaptation + stop rule vs the oracle formula
    ctl_d = t(np.concatenate([[sigma, 0, -1e300, 0, 0, 0, 0, 0, 0], np.zeros(7), mu]))
    _lib.call("pmc_mcmc_finalize", 2, _lib.ptr(ctl_d), _lib.ptr(parts), 0, None, 0, 3, 50, n, d)
    c = ctl_d.cpu().numpy()
    cap = 2.38 / d ** 0.5
    np.testing.assert_allclose(c[5], alpha_ref.mean(), rtol=1e-13)
    np.testing.assert_allclose(c[0], abs(min(sigma + 1 / 2 ** 0.75 * (alpha_ref.mean() - 0.234), min(cap, 0.99))), rtol=1e-13)
    assert c[1] == 1 and c[8] == acc.sum()


def test_device_rng_statistics():
    from pocomc_b200 import _lib
    import ctypes as C
    n, d, shape = 200000, 5, 3.5
    dev = torch.device("cuda")
    g = torch.empty(n, dtype=torch.float64, device=dev); z = torch.empty((n, d), dtype=torch.float64, device=dev)
    r = torch.empty(n, dtype=torch.float64, device=dev)
    _lib.call("pmc_rng_fill", C.c_uint64(42), C.c_uint64(1), 0, shape, _lib.ptr(g), _lib.ptr(z), _lib.ptr(r), n, d)
    g, z, r = g.cpu().numpy(), z.cpu().numpy(), r.cpu().numpy()
    assert abs(z.mean()) < 0.01 and abs(z.std() - 1) < 0.01 and abs(np.corrcoef(z[:, 0], z[:, 1])[0, 1]) < 0.01
    assert abs(r.mean() - 0.5) < 0.005 and r.min() > 0 and r.max() < 1
    assert abs(g.mean() - shape) < 0.03 and abs(g.var() - shape) < 0.1
    # a shard starting at particle offset k reproduces rows k.. of the full draw (GPU-count independent)
    z2 = torch.empty((1000, d), dtype=torch.float64, device=dev)
    _lib.call("pmc_rng_fill", C.c_uint64(42), C.c_uint64(1), 5000, 0.0, None, _lib.ptr(z2), None, 1000, d)
    np.testing.assert_array_equal(z2.cpu().numpy(), z[5000:6000])


@pytest.mark.parametrize("n,d,f32", [(1000, 64, True), (4099, 100, True), (777, 200, False), (33, 70, True)])
def test_tiled_tpcn_proposal_is_bit_identical_to_the_row_kernel(n, d, f32, monkeypatch):
    """mcmc.py:77-85 for wide problems: the 32-rows-per-block proposal kernel (csrc/mcmc_ops.cu: tpcn_propose_tiled_kernel,
    D >= 64) against the warp-per-row kernel it replaces there -- same fma chains, same reduction trees: every output bit
    for bit -- and against the numpy oracle at 1e-12.  Ragged last block, several blocks per CTA slot, f32 and f64 positions."""
    from pocomc_b200 import _lib
    rng = np.random.default_rng(n + d)
    a = rng.normal(size=(d, d)) / np.sqrt(d)
    cov = a @ a.T + 0.1 * np.eye(d)
    inv, chol = np.linalg.inv(cov), np.linalg.cholesky(cov)
    mu = rng.normal(size=d)
    theta = (rng.normal(size=(n, d)) @ chol.T + mu).astype(np.float32 if f32 else np.float64)
    zz, gg = rng.normal(size=(n, d)), rng.gamma((d + 5.0) / 2.0, size=n)
    nu, sigma = 5.0, 0.21
    dev = torch.device("cuda")
    t = lambda arr, dt=torch.float64: torch.as_tensor(np.ascontiguousarray(arr), dtype=dt).to(dev)
    ctl = np.zeros(16 + d); ctl[0] = sigma; ctl[16:] = mu
    ctl_d, th_d, g_d, z_d = t(ctl), t(theta, torch.float32 if f32 else torch.float64), t(gg), t(zz)
    inv_d, chol_d = t(inv.T), t(chol.T)
    outs = []
    for force_rows in ("1", "0"):
        monkeypatch.setenv("PMC_TPCN_ROW_KERNEL", force_rows)
        prop = torch.zeros((n, d), dtype=torch.float64, device=dev)
        prop32 = torch.zeros((n, d), dtype=torch.float32, device=dev)
        m_cur = torch.zeros(n, dtype=torch.float64, device=dev); m_prop = torch.zeros_like(m_cur)
        _lib.call("pmc_tpcn_propose", 1 if f32 else 0, _lib.ptr(th_d), _lib.ptr(ctl_d), _lib.ptr(inv_d), _lib.ptr(chol_d), nu,
                  _lib.ptr(g_d), _lib.ptr(z_d), _lib.ptr(prop), _lib.ptr(prop32), _lib.ptr(m_cur), _lib.ptr(m_prop), n, d)
        outs.append([v.cpu().numpy() for v in (prop, prop32, m_cur, m_prop)])
    for a_, b_ in zip(*outs):
        np.testing.assert_array_equal(a_, b_)
    prop_ref, m_ref = O.tpcn_propose(theta.astype(np.float64), mu, inv, chol, nu, sigma, gg, zz)
    np.testing.assert_allclose(outs[1][0], prop_ref, rtol=1e-11, atol=1e-11)
    np.testing.assert_allclose(outs[1][2], m_ref, rtol=1e-11)
    np.testing.assert_allclose(outs[1][3], O.mahalanobis(prop_ref - mu, inv), rtol=1e-11)
