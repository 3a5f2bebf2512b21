"""Proposal geometry (mu, Sigma, nu) of the MCMC kernels: the reference's ``pocomc.geometry.Geometry``
(pocomc/geometry.py:31-59) and its multivariate Student-t EM fit (pocomc/student.py:5-85).

SURVEY section 8 marks this row "next (f1)": it is adjacent to the hot path, runs once per
temperature level on [M, D] with M ~ 2 n_effective, and produces D + D^2 numbers.  The O(M) index
work (systematic resampling) runs on the GPU through pocomc_b200.tools; the small dense algebra
(weighted covariance, medians, EM scalars with digamma / bisection) is host numpy like the
reference's until the device SYRK/median kernels of f1 land."""
import numpy as np
from scipy import optimize, special

from .tools import systematic_resample

__all__ = ["Geometry", "fit_mvstud"]


def _nu_update(delta, dim, n):
    """One degrees-of-freedom update of the t EM (student.py:42-51), including the reference's
    behaviour of returning inf as soon as the score at nu = 1e300 is non-negative."""
    def score(nu):
        w = (nu + dim) / (nu + delta)
        return (-special.psi(nu / 2) + np.log(nu / 2) + np.sum(np.log(w)) / n - np.sum(w) / n + 1
                + special.psi((nu + dim) / 2) - np.log((nu + dim) / 2))
    if score(1e300) >= 0:
        return np.inf
    return optimize.bisect(score, 1e-300, 1e300)


def _delta_cannot_matter(diffs, sigma, bound=1e280):
    """True when every Mahalanobis distance delta_i = d_i^T Sigma^-1 d_i is provably finite and below ``bound``.

    The reference evaluates the nu score at nu = 1e300 first (student.py:42-51): there w = (nu + dim) / (nu + delta) is
    exactly 1.0 in floating point for any delta below ~1e284, the score is exactly 0 >= 0, and the fit returns
    nu = inf with the initial (mu, Sigma) -- whatever delta is.  delta_i <= |d_i|^2 / lambda_min(Sigma),
    so one pass over the cloud and a D x D eigenvalue problem replace the [D, D] \\ [D, n] solve (the largest single
    cost of the fit at n ~ 5e4) whenever the bound holds; anything else takes the reference's path."""
    if not (np.all(np.isfinite(sigma)) and np.allclose(sigma, sigma.T, rtol=1e-12, atol=0.0)):
        return False
    lam = np.linalg.eigvalsh(sigma)[0]
    r2 = np.max(np.einsum("ij,ij->j", diffs, diffs))
    return bool(np.isfinite(r2) and lam > 0.0 and r2 / lam < bound)


def fit_mvstud(data, tolerance=1e-6, max_iter=100):
    """EM fit of a multivariate Student-t to ``data`` [n, dim] -> (mu [dim], Sigma [dim,dim], nu)."""
    cols = np.asarray(data).T
    dim, n = cols.shape
    mu = np.median(np.ascontiguousarray(cols), axis=1)[:, None]     # same values as on the strided view, half the time
    sigma = np.cov(cols) * (n - 1) / n + np.diag(np.var(cols, axis=1)) / n
    nu, last_nu, it = 20, 0, 0
    while np.abs(last_nu - nu) > tolerance and it < max_iter:
        it += 1
        diffs = cols - mu
        if it == 1 and _delta_cannot_matter(diffs, sigma):
            return mu.T[0], sigma, np.inf
        delta = np.sum(diffs * np.linalg.solve(sigma, diffs), 0)
        last_nu = nu
        nu = _nu_update(delta, dim, n)
        if nu == np.inf:
            return mu.T[0], sigma, nu
        w = (nu + dim) / (nu + delta)
        sigma = np.dot(w * diffs, diffs.T) / n
        mu = (np.sum(w * cols, 1) / sum(w))[:, None]
    if it == max_iter:
        print("Warning: EM algorithm did not converge.")
        print("Last nu: ", last_nu)
        print("Current nu: ", nu)
    return mu.T[0], sigma, nu


class Geometry:
    """Normal (mean, cov) and Student-t (mean, cov, nu) summaries of a weighted particle cloud."""

    def __init__(self):
        self.normal_mean = None
        self.normal_cov = None
        self.t_mean = None
        self.t_cov = None
        self.t_nu = None

    def fit(self, theta, weights=None):
        """geometry.py:31-59.  With weights the t fit runs on a systematic resample of the cloud
        (one uniform from the global np.random stream)."""
        if weights is None:
            self.normal_mean = np.mean(theta, axis=0)
            self.normal_cov = np.cov(theta.T)
            cloud = theta
        else:
            self.normal_mean = np.average(theta, axis=0, weights=weights)
            self.normal_cov = np.cov(theta.T, aweights=weights)
            cloud = theta[systematic_resample(len(theta), weights=weights)]
        self.t_mean, self.t_cov, self.t_nu = fit_mvstud(cloud)
        if ~np.isfinite(self.t_nu):
            self.t_nu = 1e6
