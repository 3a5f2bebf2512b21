"""Synthetic vectorised likelihoods of the benchmark configurations and the
device fast path for scipy ``norm`` / ``uniform`` product priors (section 8 f3).

Each likelihood is an ordinary host callable ``f(x[n, D]) -> logl[n]`` (numpy, what the reference
Sampler calls) and additionally exposes ``device(x, finite, out)`` which evaluates the same
function with libpmc_b200 on CUDA f64 tensors without leaving the GPU."""
from __future__ import annotations

import math

import numpy as np
import torch

from . import _lib

LIKE_GAUSS, LIKE_ROSENBROCK, LIKE_MIXTURE, LIKE_FUNNEL = 0, 1, 2, 3

__all__ = ["CorrelatedGaussian", "Rosenbrock", "GaussianMixture", "Funnel", "DevicePrior"]


class _DeviceLikelihood:
    which = None
    p0 = 0.0
    p1 = 0.0
    mat_t = None          # numpy [D, D] (transposed precision) or None

    def _mat(self, dev):
        if self.mat_t is None:
            return None
        cache = self.__dict__.setdefault("_mat_dev", {})
        if dev not in cache:
            cache[dev] = torch.as_tensor(np.ascontiguousarray(self.mat_t, dtype=np.float64)).to(dev)
        return cache[dev]

    def __getstate__(self):
        st = self.__dict__.copy()
        st.pop("_mat_dev", None)
        return st

    def device(self, x: torch.Tensor, finite: torch.Tensor, out: torch.Tensor):
        n, d = x.shape
        _lib.call("pmc_loglike", self.which, _lib.ptr(x), _lib.ptr(finite), _lib.ptr(self._mat(x.device)),
                  float(self.p0), float(self.p1), _lib.ptr(out), n, d)


class CorrelatedGaussian(_DeviceLikelihood):
    """logL = -1/2 x^T C^-1 x - 1/2 (D log 2pi + log|C|), C = rho 11^T + (1-rho) I
    (docs/source/likelihood.ipynb pattern; BASELINE config 2)."""
    which = LIKE_GAUSS

    def __init__(self, n_dim, rho=0.95):
        self.n_dim = n_dim
        self.cov = rho * np.ones((n_dim, n_dim)) + (1 - rho) * np.eye(n_dim)
        self.prec = np.linalg.inv(self.cov)
        self.mat_t = self.prec.T.copy()
        self.p0 = -0.5 * (n_dim * math.log(2 * math.pi) + np.linalg.slogdet(self.cov)[1])

    def __call__(self, x):
        return -0.5 * np.einsum("ij,ij->i", x @ self.prec, x) + self.p0        # one GEMM + a row-wise dot (a 3-operand einsum is ~40x slower)

    def analytic_logz(self, prior_sd):
        c = self.cov + prior_sd ** 2 * np.eye(self.n_dim)
        return -0.5 * (self.n_dim * math.log(2 * math.pi) + np.linalg.slogdet(c)[1])


class Rosenbrock(_DeviceLikelihood):
    """README.md:53-55: -sum_{i even} 10 (x_i^2 - x_{i+1})^2 + (x_i - 1)^2."""
    which = LIKE_ROSENBROCK

    def __call__(self, x):
        return -np.sum(10.0 * (x[:, ::2] ** 2.0 - x[:, 1::2]) ** 2.0 + (x[:, ::2] - 1.0) ** 2.0, axis=1)


class GaussianMixture(_DeviceLikelihood):
    """logaddexp(N(x; +c 1, s^2 I), N(x; -c 1, s^2 I)) - log 2 (BASELINE config 3)."""
    which = LIKE_MIXTURE

    def __init__(self, c=1.5, s=0.5):
        self.p0, self.p1 = c, s

    def __call__(self, x):
        d = x.shape[1]
        norm = -0.5 * d * math.log(2 * math.pi * self.p1 ** 2)
        a = norm - 0.5 * np.sum((x - self.p0) ** 2, axis=1) / self.p1 ** 2
        b = norm - 0.5 * np.sum((x + self.p0) ** 2, axis=1) / self.p1 ** 2
        return np.logaddexp(a, b) - math.log(2.0)


class Funnel(_DeviceLikelihood):
    """Neal's funnel: x0 ~ N(0, sd^2), x_i ~ N(0, exp(x0)) (BASELINE config 5)."""
    which = LIKE_FUNNEL

    def __init__(self, sd=3.0):
        self.p0 = sd

    def __call__(self, x):
        x0 = x[:, 0]
        d = x.shape[1]
        return (-0.5 * x0 ** 2 / self.p0 ** 2 - 0.5 * math.log(2 * math.pi * self.p0 ** 2)
                - 0.5 * np.sum(x[:, 1:] ** 2, axis=1) * np.exp(-x0) - 0.5 * (d - 1) * (math.log(2 * math.pi) + x0))


class DevicePrior:
    """logpdf of a product of norm(loc, scale) / uniform(loc, loc+scale) factors on CUDA tensors;
    also clears ``finite`` where the prior is not finite (mcmc.py:108-109)."""

    def __init__(self, kind, loc, scale):
        self.kind, self.loc, self.scale = (np.asarray(kind, np.int32), np.asarray(loc, np.float64),
                                           np.asarray(scale, np.float64))
        self._dev = {}

    def _params(self, dev):
        if dev not in self._dev:
            self._dev[dev] = tuple(torch.from_numpy(a.copy()).to(dev) for a in (self.kind, self.loc, self.scale))
        return self._dev[dev]

    def __call__(self, x: torch.Tensor, finite: torch.Tensor, out: torch.Tensor):
        kind, loc, scale = self._params(x.device)
        n, d = x.shape
        _lib.call("pmc_logprior", _lib.ptr(x), _lib.ptr(finite), _lib.ptr(kind), _lib.ptr(loc), _lib.ptr(scale),
                  _lib.ptr(out), n, d)

    def bind(self, x: torch.Tensor, finite: torch.Tensor, out: torch.Tensor):
        """Pre-bound variant of ``__call__`` for fixed buffers (the per-step loop of the MCMC engine)."""
        kind, loc, scale = self._params(x.device)
        n, d = x.shape
        return _lib.bind("pmc_logprior", _lib.ptr(x), _lib.ptr(finite), _lib.ptr(kind), _lib.ptr(loc), _lib.ptr(scale),
                         _lib.ptr(out), n, d)
