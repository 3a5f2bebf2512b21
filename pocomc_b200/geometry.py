"""Proposal geometry (mu, Sigma, nu) of the MCMC kernels: the reference's ``pocomc.geometry.Geometry``
(pocomc/geometry.py:31-59) and its multivariate Student-t EM fit (pocomc/student.py:5-85).

SURVEY section 8 (f1): adjacent to the hot path, once per temperature level on [M, D] with M ~ 2 n_effective (up to 10^6 x
200 at the BASELINE sizes), producing D + D^2 numbers.  Everything that touches the cloud runs on the GPU -- weighted column
sums, the centred weighted scatter matrix (np.cov numerators and the EM update), per-dimension medians, Mahalanobis
distances, the sums behind the degrees-of-freedom score (csrc/geom_ops.cu: f64, fixed-order reductions), the systematic
resample and the row gather (csrc/smc_ops.cu); the D x D algebra and the scalar root finding (digamma, bisection) stay on
the host exactly like the reference's.  ``fit_mvstud_host`` / ``Geometry(host=True)`` is the reference's numpy formulation
kept next to it: the CPU test-suite pins it bit for bit against recorded reference vectors and the GPU tests compare the
device fit with it at 1e-12."""
import numpy as np
from scipy import optimize, special

import torch

from . import _lib
from .tools import systematic_resample

__all__ = ["Geometry", "fit_mvstud", "fit_mvstud_host", "fit_mvstud_device"]


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


def _nu_update_sums(score_sums, dim, n):
    """_nu_update with the two sums over the cloud supplied by a callable nu -> (sum log w, sum w) (device reductions)"""
    def score(nu):
        sl, sw = score_sums(nu)
        return (-special.psi(nu / 2) + np.log(nu / 2) + sl / n - sw / n + 1
                + special.psi((nu + dim) / 2) - np.log((nu + dim) / 2))
    if score(1e300) >= 0:
        return np.inf
    return optimize.bisect(score, 1e-300, 1e300)


class _Cloud:
    """an [n, d] f64 cloud resident on the GPU and the reductions over it (csrc/geom_ops.cu)"""

    def __init__(self, x):
        if not torch.is_tensor(x):
            x = torch.from_numpy(np.ascontiguousarray(x, dtype=np.float64))
        self.x = x.to(_lib.device(), torch.float64).contiguous()
        self.n, self.d = int(self.x.shape[0]), int(self.x.shape[1])
        self.dev = self.x.device
        self.scratch = torch.empty(int(_lib.load().pmc_geometry_scratch_size(self.n, self.d)), dtype=torch.float64, device=self.dev)

    def _vec(self, v):
        return None if v is None else torch.as_tensor(np.ascontiguousarray(v, dtype=np.float64)).to(self.dev)

    def colsums(self, w=None, center=None):
        """(sum_i w_i x_i [d], sum w, sum w^2, max_i |x_i - center|^2)"""
        c = self._vec(center)
        out = torch.empty(self.d + 3, dtype=torch.float64, device=self.dev)
        _lib.call("pmc_weighted_colsums", _lib.ptr(self.x), _lib.ptr(w), _lib.ptr(c), self.n, self.d, _lib.ptr(self.scratch), _lib.ptr(out))
        o = out.cpu().numpy()
        return o[:self.d].copy(), float(o[self.d]), float(o[self.d + 1]), float(o[self.d + 2])

    def scatter(self, center, w=None):
        """sum_i w_i (x_i - c)(x_i - c)^T [d, d]"""
        c = self._vec(center)
        out = torch.empty((self.d, self.d), dtype=torch.float64, device=self.dev)
        _lib.call("pmc_weighted_scatter", _lib.ptr(self.x), _lib.ptr(w), _lib.ptr(c), self.n, self.d, _lib.ptr(self.scratch), _lib.ptr(out))
        return out.cpu().numpy()

    def medians(self):
        """np.median(x, axis=0): per-dimension sort on the device, mean of the two middle order statistics"""
        s, _ = torch.sort(self.x.t().contiguous(), dim=1)
        lo, hi = (self.n - 1) // 2, self.n // 2
        return ((s[:, lo] + s[:, hi]) / 2.0).cpu().numpy() if lo != hi else s[:, lo].cpu().numpy()

    def mahalanobis(self, center, precision):
        c, p = self._vec(center), self._vec(precision)
        delta = torch.empty(self.n, dtype=torch.float64, device=self.dev)
        _lib.call("pmc_mahalanobis", _lib.ptr(self.x), _lib.ptr(c), _lib.ptr(p), self.n, self.d, _lib.ptr(delta))
        return delta

    def student_weights(self, delta, nu, store=False):
        w = torch.empty(self.n, dtype=torch.float64, device=self.dev) if store else None
        out = torch.empty(2, dtype=torch.float64, device=self.dev)
        _lib.call("pmc_student_weights", _lib.ptr(delta), self.n, float(nu), float(self.d), _lib.ptr(w), _lib.ptr(self.scratch), _lib.ptr(out))
        o = out.cpu().numpy()
        return float(o[0]), float(o[1]), w

    def take(self, idx):
        from .tools import gather_rows_device
        return _Cloud(gather_rows_device(self.x, idx))


def fit_mvstud_device(cloud, tolerance=1e-6, max_iter=100):
    """student.py:5-85 with every pass over the cloud on the GPU; same iteration, same exits, same quirks as the host
    formulation below (the nu = 1e300 score test ends the fit after the initial moments whenever the Mahalanobis
    distances are provably below 1e280 -- then they are never computed)."""
    if not isinstance(cloud, _Cloud):
        cloud = _Cloud(cloud)
    n, dim = cloud.n, cloud.d
    mu = cloud.medians()
    sums, _, _, r2max = cloud.colsums(center=mu)
    mean = sums / n
    s0 = cloud.scatter(mean) / n                                   # = np.cov(cols) (n - 1) / n
    sigma = s0 + np.diag(np.diag(s0)) / n                          # + diag(np.var(cols, axis=1)) / n
    nu, last_nu, it = 20, 0, 0
    while np.abs(last_nu - nu) > tolerance and it < max_iter:
        it += 1
        if it == 1 and np.all(np.isfinite(sigma)) and np.allclose(sigma, sigma.T, rtol=1e-12, atol=0.0):
            lam = np.linalg.eigvalsh(sigma)[0]
            if np.isfinite(r2max) and lam > 0.0 and r2max / lam < 1e280:
                return mu, sigma, np.inf
        delta = cloud.mahalanobis(mu, np.linalg.inv(sigma))
        last_nu = nu
        nu = _nu_update_sums(lambda v: cloud.student_weights(delta, v)[:2], dim, n)
        if nu == np.inf:
            return mu, sigma, nu
        _, sw, w = cloud.student_weights(delta, nu, store=True)
        sigma = cloud.scatter(mu, w) / n
        mu = cloud.colsums(w)[0] / sw
    if it == max_iter:
        print("Warning: EM algorithm did not converge.")
        print("Last nu: ", last_nu)
        print("Current nu: ", nu)
    return mu, sigma, nu


def fit_mvstud(data, tolerance=1e-6, max_iter=100):
    """EM fit of a multivariate Student-t to ``data`` [n, dim] -> (mu [dim], Sigma [dim,dim], nu) (student.py:5-85), on the GPU."""
    return fit_mvstud_device(data, tolerance, max_iter)


def fit_mvstud_host(data, tolerance=1e-6, max_iter=100):
    """The reference's numpy formulation of the same fit (student.py:5-85), bit for bit: what the device fit is checked against."""
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

    def __init__(self, host=False):
        self.host = bool(host)      # True: the reference's numpy formulation (CPU test-suite); default: GPU reductions
        self.normal_mean = None
        self.normal_cov = None
        self.t_mean = None
        self.t_cov = None
        self.t_nu = None

    def fit(self, theta, weights=None):
        """geometry.py:31-59.  With weights the t fit runs on a systematic resample of the cloud
        (one uniform from the global np.random stream)."""
        if not self.host:
            return self._fit_device(theta, weights)
        if weights is None:
            self.normal_mean = np.mean(theta, axis=0)
            self.normal_cov = np.cov(theta.T)
            cloud = theta
        else:
            self.normal_mean = np.average(theta, axis=0, weights=weights)
            self.normal_cov = np.cov(theta.T, aweights=weights)
            cloud = theta[systematic_resample(len(theta), weights=weights)]
        self.t_mean, self.t_cov, self.t_nu = fit_mvstud_host(cloud)
        if ~np.isfinite(self.t_nu):
            self.t_nu = 1e6

    def _fit_device(self, theta, weights=None):
        """the same fit with the cloud resident on the GPU (one upload of theta / weights, D + D^2 numbers back per moment)"""
        cloud = _Cloud(theta)
        n = cloud.n
        if weights is None:
            sums, _, _, _ = cloud.colsums()
            self.normal_mean = sums / n
            self.normal_cov = cloud.scatter(self.normal_mean) / (n - 1)                     # np.cov(theta.T)
            sub = cloud
        else:
            wh = np.ascontiguousarray(weights, dtype=np.float64)
            w = torch.from_numpy(wh).to(cloud.dev)
            sums, sw, sw2, _ = cloud.colsums(w)
            self.normal_mean = sums / sw                                                     # np.average(theta, 0, weights)
            self.normal_cov = cloud.scatter(self.normal_mean, w) / (sw - sw2 / sw)           # np.cov(theta.T, aweights=w)
            idx = systematic_resample(n, weights=wh)                                         # one uniform from np.random
            sub = cloud.take(torch.from_numpy(np.ascontiguousarray(idx, dtype=np.int64)).to(cloud.dev))
        self.t_mean, self.t_cov, self.t_nu = fit_mvstud_device(sub)
        if ~np.isfinite(self.t_nu):
            self.t_nu = 1e6
