"""pc.Prior and pc.geometry.Geometry: the reference's own prior tests (tests/test_prior.py:10-51)
restated against pocomc_b200, the device fast-path descriptor, and the proposal-geometry fit
(pocomc/geometry.py:31-59, pocomc/student.py:5-85) against vectors recorded from the reference
(tests/golden/geometry.npz, oracle/make_golden.py)."""
import numpy as np
import pytest
from scipy.stats import halfnorm, norm, uniform

import pocomc_b200 as pc

F64 = dict(rtol=1e-10, atol=1e-12)
KEYS = ("normal_mean", "normal_cov", "t_mean", "t_cov", "t_nu")


# ---- reference tests/test_prior.py ------------------------------------------------------------
def test_prior_sample_shape():
    prior = pc.Prior([norm(0, 1), norm(0, 1)])
    assert np.shape(prior.rvs(10)) == (10, 2)


def test_prior_logpdf_like_reference():
    prior = pc.Prior([norm(0, 1), norm(0, 1)])
    x = prior.rvs(10)
    lp = prior.logpdf(x)
    assert isinstance(lp, np.ndarray)
    assert lp.shape == (10,)
    assert np.all(lp < 0)
    assert np.all(np.isfinite(lp))
    np.testing.assert_allclose(lp, norm(0, 1).logpdf(x).sum(axis=1), rtol=1e-14)


def test_prior_bounds_and_dim():
    prior = pc.Prior([norm(0, 1), uniform(-2.0, 5.0), halfnorm(0.0, 2.0)])
    b = prior.bounds
    assert b.shape == (3, 2)
    assert np.all(b[:, 0] < b[:, 1])
    np.testing.assert_array_equal(b, [[-np.inf, np.inf], [-2.0, 3.0], [0.0, np.inf]])
    assert prior.dim == 3


def test_prior_rvs_consumes_the_global_stream_in_dimension_order():
    """prior.py:102-132: one dist.rvs(size) per dimension, in order, from np.random's global state."""
    dists = [norm(1.0, 2.0), uniform(-1.0, 2.0)]
    np.random.seed(5)
    got = pc.Prior(dists).rvs(7)
    np.random.seed(5)
    want = np.transpose([d.rvs(size=7) for d in dists])
    np.testing.assert_array_equal(got, want)


def test_device_spec_only_for_norm_and_uniform_factors():
    kind, loc, scale = pc.Prior([norm(1.0, 2.0), uniform(-1.0, 4.0), norm(), uniform(loc=3.0, scale=0.5)]).device_spec()
    np.testing.assert_array_equal(kind, [0, 1, 0, 1])
    np.testing.assert_array_equal(loc, [1.0, -1.0, 0.0, 3.0])
    np.testing.assert_array_equal(scale, [2.0, 4.0, 1.0, 0.5])
    assert pc.Prior([norm(0, 1), halfnorm()]).device_spec() is None        # anything else stays a host black box
    assert pc.Prior([norm(np.zeros(2), 1.0)]).device_spec() is None        # array-valued parameters too


# ---- geometry -----------------------------------------------------------------------------------
def test_unweighted_geometry_matches_reference(golden):
    g = golden("geometry")
    geo = pc.geometry.Geometry(host=True)
    geo.fit(g["theta"])
    for k in KEYS:
        np.testing.assert_allclose(getattr(geo, k), g["nw_" + k], err_msg=k, **F64)


def test_fit_mvstud_keeps_the_reference_nu_quirk():
    """student.py:42-51 evaluates the nu score at 1e300 first and returns nu = inf when it is
    non-negative -- which, in floating point, it is even for a genuinely heavy-tailed cloud -- and
    Geometry.fit then pins nu to 1e6 (geometry.py:57-58).  Parity mode keeps that behaviour."""
    rng = np.random.default_rng(4)                       # multivariate t_3: normal / sqrt(chi2_3 / 3)
    heavy = rng.normal(size=(4000, 5)) / np.sqrt(rng.chisquare(3.0, size=(4000, 1)) / 3.0)
    mu, cov, nu = pc.geometry.fit_mvstud_host(heavy)
    assert mu.shape == (5,) and cov.shape == (5, 5)
    assert nu == np.inf
    np.testing.assert_allclose(mu, np.median(heavy, axis=0), rtol=1e-14)          # first-iteration exit: the initial guess
    geo = pc.geometry.Geometry(host=True)
    geo.fit(heavy)
    assert geo.t_nu == 1e6


@pytest.mark.gpu
@pytest.mark.parametrize("host", [True, False])
def test_weighted_geometry_matches_reference(golden, host):
    """Weighted fit against the reference's recorded vectors: the systematic resample (one uniform from the global stream,
    tools.py:136-186) runs on the GPU -- indices must be bit-exact for the t fit to agree -- and with host=False so does
    every pass over the cloud (csrc/geom_ops.cu): moments, medians, scatter; f64 bar 1e-12."""
    g = golden("geometry")
    geo = pc.geometry.Geometry(host=host)
    np.random.seed(77)
    geo.fit(g["theta"], weights=g["w"])
    for k in KEYS:
        np.testing.assert_allclose(getattr(geo, k), g[k], err_msg=k, **F64)


@pytest.mark.gpu
def test_device_geometry_matches_reference_unweighted_and_mvstud(golden):
    g = golden("geometry")
    geo = pc.geometry.Geometry()
    geo.fit(g["theta"])
    for k in KEYS:
        np.testing.assert_allclose(getattr(geo, k), g["nw_" + k], err_msg=k, **F64)
    mu, sigma, nu = pc.geometry.fit_mvstud(g["theta"])
    np.testing.assert_allclose(mu, g["mvstud_mu"], **F64)
    np.testing.assert_allclose(sigma, g["mvstud_sigma"], **F64)
    assert nu == g["mvstud_nu"] or (np.isinf(nu) and np.isinf(g["mvstud_nu"]))


@pytest.mark.gpu
@pytest.mark.parametrize("n,d", [(777, 3), (5000, 32), (20011, 200), (300, 70)])
def test_device_geometry_reductions_against_numpy(n, d):
    """every reduction of csrc/geom_ops.cu against its numpy formula (student.py:38-58, geometry.py:44-49): odd sizes, several
    row chunks, several 64 x 64 output tiles, weights with a large dynamic range; f64 bar 1e-12 relative to the result's scale."""
    import torch
    from pocomc_b200.geometry import _Cloud, fit_mvstud_device, fit_mvstud_host
    rng = np.random.default_rng(n + d)
    a = rng.normal(size=(d, d)) / np.sqrt(d)
    x = rng.normal(size=(n, d)) @ a + rng.normal(size=d)
    w = np.exp(rng.normal(size=n) * 3.0)
    cloud = _Cloud(x)
    wd = torch.from_numpy(w).cuda()
    mu = np.median(x, axis=0)
    np.testing.assert_array_equal(cloud.medians(), mu)
    sums, sw, sw2, r2 = cloud.colsums(wd, center=mu)
    np.testing.assert_allclose(sums, w @ x, rtol=1e-12, atol=1e-12 * np.abs(w @ x).max())
    np.testing.assert_allclose([sw, sw2, r2], [w.sum(), (w * w).sum(), ((x - mu) ** 2).sum(1).max()], rtol=1e-12)
    ref_cov = np.cov(x.T, aweights=w)
    got_cov = cloud.scatter(np.average(x, axis=0, weights=w), wd) / (sw - sw2 / sw)
    np.testing.assert_allclose(got_cov, ref_cov, rtol=1e-11, atol=1e-12 * np.abs(ref_cov).max())
    assert np.array_equal(got_cov, got_cov.T)
    sigma = np.cov(x.T)
    diffs = (x - mu).T
    ref_delta = np.sum(diffs * np.linalg.solve(sigma, diffs), 0)
    delta = cloud.mahalanobis(mu, np.linalg.inv(sigma))
    np.testing.assert_allclose(delta.cpu().numpy(), ref_delta, rtol=1e-9)
    for nu in (0.7, 20.0, 1e300):
        sl, s1, wt = cloud.student_weights(delta, nu, store=True)
        wr = (nu + d) / (nu + delta.cpu().numpy())
        np.testing.assert_allclose([sl, s1], [np.log(wr).sum(), wr.sum()], rtol=1e-12, atol=1e-9)
        np.testing.assert_allclose(wt.cpu().numpy(), wr, rtol=1e-14)
    # the whole fit: same exit, same values as the reference's numpy formulation
    md, sd, nd = fit_mvstud_device(x)
    mh, sh_, nh = fit_mvstud_host(x)
    np.testing.assert_allclose(md, mh, rtol=1e-12)
    np.testing.assert_allclose(sd, sh_, rtol=1e-11, atol=1e-12 * np.abs(sh_).max())
    assert nd == nh or (np.isinf(nd) and np.isinf(nh))


def test_fit_mvstud_matches_reference_vectors_and_takes_the_full_path_when_it_must(golden):
    """student.py:5-85 on the recorded cloud (bit-for-bit: the shortcut past the unused Mahalanobis solve must not
    change a single value), and the guard that sends degenerate clouds down the reference's own path."""
    g = golden("geometry")
    mu, sigma, nu = pc.geometry.fit_mvstud_host(g["theta"])
    np.testing.assert_array_equal(mu, g["mvstud_mu"])
    np.testing.assert_array_equal(sigma, g["mvstud_sigma"])
    assert nu == g["mvstud_nu"] or (np.isinf(nu) and np.isinf(g["mvstud_nu"]))
    from pocomc_b200.geometry import _delta_cannot_matter
    rng = np.random.default_rng(1)
    x = rng.normal(size=(200, 3))
    diffs = (x - np.median(x, axis=0)).T
    assert _delta_cannot_matter(diffs, np.cov(x.T))
    dup = np.c_[x, x[:, 0]]                                         # singular covariance: no bound on delta
    assert not _delta_cannot_matter((dup - np.median(dup, axis=0)).T, np.cov(dup.T) * 0.0)
    bad = diffs.copy(); bad[0, 0] = np.inf
    assert not _delta_cannot_matter(bad, np.cov(x.T))
    assert not _delta_cannot_matter(diffs * 1e150, np.cov(x.T))       # |d|^2 / lambda_min overflows the bound


def test_particles_take_flat_equals_concatenate_then_index():
    """Sampler._reweight's trimmed gather (sampler.py:792-800): same rows, same dtype, any index pattern."""
    from pocomc_b200.particles import Particles
    rng = np.random.default_rng(2)
    p = Particles(7, 3)
    for t in range(9):
        p.update(dict(u=rng.normal(size=(7, 3)), logl=rng.normal(size=7), iter=t))
    for idx in (np.array([0, 1, 6, 7, 20, 62]), np.arange(63), np.array([], dtype=np.int64), np.array([62]),
                np.array([5, 5, 5, 40]), np.array([40, 3, 9])):                      # the last one is unsorted: fallback
        for key in ("u", "logl"):
            want = p.get(key, flat=True)[idx]
            got = p.take_flat(key, idx)
            assert got.dtype == want.dtype and got.shape == want.shape
            np.testing.assert_array_equal(got, want)
    with pytest.raises(IndexError):
        p.take_flat("u", np.array([0, 63]))
