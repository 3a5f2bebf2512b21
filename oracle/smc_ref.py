"""TEST INFRASTRUCTURE -- numpy restatement of pocoMC's SMC/MCMC hot path (CPU oracle).

Every function cites the reference lines it restates (paths relative to /root/reference).
Pinned against the reference itself: oracle/make_golden.py imports the reference modules in the
build container, records inputs / pre-drawn noise / outputs under tests/golden/, and
tests/test_oracle_golden.py checks this file against those vectors and against the reference's
known answers.  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import this module; the product never does.

Conventions
-----------
* SMC state (u, x, logdetj, logl, logp) is float64; flow outputs (theta, logdetj_flow) are
  float32 exactly as they come back from the reference's numpy<->torch shim (tools.py:276,292).
* All randomness is passed in explicitly (``Noise``) in the order the reference consumes it from
  the legacy global ``np.random`` stream: per tpCN step N standard gammas, then
  N*D normals, then (after the likelihood) N uniforms.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Callable, Optional

import numpy as np
from scipy.special import erf, erfinv

SQRTEPS = math.sqrt(float(np.finfo(np.float64).eps))

# --------------------------------------------------------------------------------------
# Reparameterisation ("scaler")  -- pocomc/scaler.py
# --------------------------------------------------------------------------------------
KIND_NONE, KIND_LEFT, KIND_RIGHT, KIND_BOTH = 0, 1, 2, 3


@dataclass
class ScalerParams:
    """Flat description of a fitted Reparameterize (scaler.py:52-73,159-178)."""
    low: np.ndarray      # [D] f64, non-finite = unbounded
    high: np.ndarray     # [D]
    mu: np.ndarray       # [D]
    sigma: np.ndarray    # [D]
    logit: bool = False  # transform == "logit" (else probit)
    scale: bool = True

    @property
    def kind(self) -> np.ndarray:
        """Per-dimension bound kind, scaler.py:459-490."""
        lo, hi = np.isfinite(self.low), np.isfinite(self.high)
        return np.where(lo & hi, KIND_BOTH, np.where(lo, KIND_LEFT, np.where(hi, KIND_RIGHT, KIND_NONE)))


def scaler_bounded_forward(x, p: ScalerParams):
    """x -> v (bounded -> unbounded, no affine); scaler.py:228-247,315-327,348-360,380-400,427-440."""
    x = np.asarray(x, dtype=np.float64)
    v = np.empty(x.shape)
    kind = p.kind
    for d in range(x.shape[1]):
        xd = x[:, d]
        if kind[d] == KIND_NONE:
            v[:, d] = xd
        elif kind[d] == KIND_LEFT:
            v[:, d] = np.log(xd - p.low[d])
        elif kind[d] == KIND_RIGHT:
            v[:, d] = np.log(p.high[d] - xd)
        else:
            q = (xd - p.low[d]) / (p.high[d] - p.low[d])   # reference never clips (scaler.py:393 discards)
            v[:, d] = np.log(q / (1.0 - q)) if p.logit else np.sqrt(2.0) * erfinv(2.0 * q - 1.0)
    return v


def scaler_fit(x, low, high, logit=False) -> ScalerParams:
    """scaler.py:159-173 (diagonal branch): mu/sigma = mean/std (ddof 0) of the bounded-forward."""
    p = ScalerParams(np.asarray(low, float), np.asarray(high, float), None, None, logit)
    v = scaler_bounded_forward(x, p)
    p.mu = np.mean(v, axis=0)
    p.sigma = np.std(v, axis=0)
    return p


def scaler_forward(x, p: ScalerParams):
    """scaler.py:180-202 + 276-291."""
    v = scaler_bounded_forward(x, p)
    return (v - p.mu) / p.sigma if p.scale else v


def scaler_inverse(u, p: ScalerParams):
    """u -> (x, log|dx/du|); scaler.py:204-226,249-274,293-313,329-346,362-378,402-425,442-457."""
    u = np.asarray(u)
    if p.scale:
        v = p.mu + p.sigma * u                      # promotes f32 u to f64 (scaler.py:310)
        logdet = np.sum(np.log(p.sigma)) * np.ones(len(u))
    else:
        v = u.astype(np.float64)
        logdet = np.zeros(len(u))
    x = np.empty(v.shape)
    J = np.empty(v.shape)
    kind = p.kind
    for d in range(v.shape[1]):
        vd = v[:, d]
        if kind[d] == KIND_NONE:
            x[:, d], J[:, d] = vd, 0.0
        elif kind[d] == KIND_LEFT:
            x[:, d], J[:, d] = np.exp(vd) + p.low[d], vd
        elif kind[d] == KIND_RIGHT:
            x[:, d], J[:, d] = p.high[d] - np.exp(vd), vd
        else:
            span = p.high[d] - p.low[d]
            if p.logit:
                q = np.exp(-np.logaddexp(0, -vd))
                J[:, d] = np.log(span) + np.log(q) + np.log(1.0 - q)
            else:
                q = (erf(vd / np.sqrt(2.0)) + 1.0) / 2.0
                J[:, d] = np.log(span) + (-vd ** 2.0 / 2.0) - np.log(np.sqrt(2.0 * np.pi))
            x[:, d] = q * span + p.low[d]
    return x, logdet + np.sum(J, axis=1)


def apply_boundary_conditions(x, p: ScalerParams, periodic=None, reflective=None):
    """Wrap periodic dims, then reflect reflective dims (scaler.py:84-157)."""
    x = np.array(x, dtype=np.float64, copy=True)
    for d in (periodic or []):
        lo, hi = p.low[d], p.high[d]
        col = x[:, d]
        for j in range(len(col)):
            while col[j] > hi:
                col[j] = lo + col[j] - hi
            while col[j] < lo:
                col[j] = hi + col[j] - lo
    for d in (reflective or []):
        lo, hi = p.low[d], p.high[d]
        col = x[:, d]
        for j in range(len(col)):
            while col[j] > hi:
                col[j] = hi - col[j] + hi
            while col[j] < lo:
                col[j] = lo + lo - col[j]
    return x


# --------------------------------------------------------------------------------------
# MCMC kernels  -- pocomc/mcmc.py
# --------------------------------------------------------------------------------------
def mahalanobis(diff, inv_cov):
    """diff_k^T inv_cov diff_k per row (mcmc.py:80,128-129)."""
    return np.einsum("ki,ki->k", diff, diff @ inv_cov.T)


def tpcn_propose(pos, mu, inv_cov, chol, nu, sigma, g, z):
    """t-preconditioned Crank-Nicolson proposal (mcmc.py:77-85 / 409-417).

    ``g`` are N standard-gamma((D+nu)/2) draws, ``z`` N x D standard normals.  The reference's
    ``np.random.gamma(a, scale_k)`` equals ``scale_k * standard_gamma(a)`` draw for draw.
    Returns (proposal f64 [N,D], m = Mahalanobis distance of the current position).
    """
    n_dim = pos.shape[1]
    diff = pos - mu                                  # f32 - f64 -> f64 for the preconditioned kernel
    m = mahalanobis(diff, inv_cov)
    s = 1.0 / (g * (2.0 / (nu + m)))
    prop = mu + (1.0 - sigma ** 2.0) ** 0.5 * diff + sigma * np.sqrt(s)[:, None] * (z @ chol.T)
    return prop, m


def rwm_propose(pos, chol, sigma, z):
    """Random-walk proposal (mcmc.py:251-253 / 569-571)."""
    return pos + sigma * (z @ chol.T)


def t_factor(m, n_dim, nu):
    """-(D+nu)/2 * log(1 + m/nu)   (mcmc.py:128-129)."""
    return -(n_dim + nu) / 2 * np.log(1 + m / nu)


def mh_alpha(beta, logl_p, logl, logp_p, logp, ldj_p, ldj, ldjf_p=None, ldjf=None, A=None, B=None):
    """Acceptance probability with the reference's left-to-right summation (mcmc.py:130-134 etc.)."""
    t = logl_p * beta - logl * beta + logp_p - logp + ldj_p - ldj
    if ldjf_p is not None:
        t = t + ldjf_p - ldjf
    if A is not None:
        t = t - A + B
    with np.errstate(over="ignore", invalid="ignore"):
        alpha = np.minimum(np.ones(len(t)), np.exp(t))
    alpha[np.isnan(alpha)] = 0.0
    return alpha


@dataclass
class Noise:
    """Explicit randomness for one MCMC step (consumption order of SURVEY App. F)."""
    g: Optional[np.ndarray]   # [N] standard gamma((D+nu)/2) -- tpCN kernels only
    z: np.ndarray             # [N, D] standard normal
    r: Optional[np.ndarray] = None   # [N] uniform, drawn AFTER the likelihood call


class GlobalNumpyNoise:
    """Draws from the legacy global np.random stream exactly as the reference's loops do."""

    def gamma_normal(self, n, d, shape):
        g = np.random.standard_gamma(shape, size=n) if shape is not None else None
        z = np.random.randn(n, d)
        return g, z

    def uniform(self, n):
        return np.random.rand(n)


class ReplayNoise:
    """Replays pre-drawn per-step noise (lists indexed by step)."""

    def __init__(self, steps):
        self.steps, self.i = list(steps), 0

    def gamma_normal(self, n, d, shape):
        s = self.steps[self.i]
        return s.g, s.z

    def uniform(self, n):
        s = self.steps[self.i]
        self.i += 1
        return s.r


def mcmc_kernel(kind: str, state: dict, loglike: Callable, logprior: Callable, scaler: ScalerParams,
                geometry: dict, options: dict, flow=None, noise=None, periodic=None, reflective=None,
                trace: Optional[list] = None):
    """One ``_mutate`` call: restates the four reference kernels as one parametrised loop.

    kind in {"tpcn_flow", "rwm_flow", "tpcn", "rwm"} = preconditioned_pcn (mcmc.py:8-183),
    preconditioned_rwm (:186-341), pcn (:344-506), rwm (:508-654).
    ``flow`` must expose forward(u)->(theta f32, -ladj f32) and inverse(theta)->(u f32, ladj f32)
    like tools.flow_numpy_wrapper (tools.py:318-349).  ``geometry``: t_mean,t_cov,t_nu | normal_cov.
    """
    use_flow = kind.endswith("_flow")
    tp = kind.startswith("tpcn")
    noise = noise or GlobalNumpyNoise()
    u = np.copy(state["u"]); x = np.copy(state["x"])
    logdetj = np.copy(state["logdetj"]); logl = np.copy(state["logl"]); logp = np.copy(state["logp"])
    beta = state["beta"]
    n, n_dim = x.shape
    n_max, n_steps = options["n_max"], options["n_steps"]
    sigma = options["proposal_scale"]
    if tp:
        sigma = np.minimum(sigma, 0.99)
        mu = np.array(geometry["t_mean"], dtype=np.float64)
        cov, nu = geometry["t_cov"], geometry["t_nu"]
        inv_cov = np.linalg.inv(cov)
        chol = np.linalg.cholesky(cov)
    else:
        chol = np.linalg.cholesky(geometry["normal_cov"])
    if use_flow:
        pos, ldjf = flow.forward(u)          # theta f32, logdetj_flow f32 (sign already flipped)
    else:
        pos, ldjf = u, None
    track = (lambda: np.mean(logl + logp)) if tp else (lambda: np.mean(logl + logp + logdetj))
    best, cnt, n_calls, i = track(), 0, 0, 0
    cap = 2.38 / n_dim ** 0.5
    while True:
        i += 1
        g, z = noise.gamma_normal(n, n_dim, (n_dim + nu) / 2 if tp else None)
        if tp:
            prop, m = tpcn_propose(pos, mu, inv_cov, chol, nu, sigma, g, z)
        else:
            prop = rwm_propose(pos, chol, sigma, z)
        if use_flow:
            u_p, ldjf_p = flow.inverse(prop)
        else:
            u_p, ldjf_p = prop, None
        x_p, ldj_p = scaler_inverse(u_p, scaler)
        if periodic is not None or reflective is not None:
            x_p = apply_boundary_conditions(x_p, scaler, periodic, reflective)
            u_p = scaler_forward(x_p, scaler)
            x_p, ldj_p = scaler_inverse(u_p, scaler)
        finite = np.isfinite(ldj_p) & np.isfinite(x_p).all(axis=1)
        logp_p = np.full(n, -np.inf)
        if finite.any():
            logp_p[finite] = logprior(x_p[finite])
        finite = finite & np.isfinite(logp_p)
        logl_p = np.full(n, -np.inf)
        if finite.any():
            logl_p[finite] = loglike(x_p[finite])
        n_calls += int(np.sum(finite))
        if tp:
            A = t_factor(mahalanobis(prop - mu, inv_cov), n_dim, nu)
            B = t_factor(m, n_dim, nu)
        else:
            A = B = None
        alpha = mh_alpha(beta, logl_p, logl, logp_p, logp, ldj_p, logdetj, ldjf_p, ldjf, A, B)
        r = noise.uniform(n)
        acc = r < alpha
        if use_flow:
            pos[acc] = prop[acc]             # rounds the f64 proposal to the f32 theta array (mcmc.py:141)
            ldjf[acc] = ldjf_p[acc]
        u[acc] = u_p[acc]; x[acc] = x_p[acc]
        logdetj[acc] = ldj_p[acc]; logl[acc] = logl_p[acc]; logp[acc] = logp_p[acc]
        mean_alpha = np.mean(alpha)
        if tp:
            sigma = np.abs(np.minimum(sigma + 1 / (i + 1) ** 0.75 * (mean_alpha - 0.234), np.minimum(cap, 0.99)))
            if use_flow:
                mu = mu + 1.0 / (i + 1.0) * (np.mean(pos, axis=0) - mu)
        elif use_flow:
            sigma = sigma + 1 / (i + 1) * (mean_alpha - 0.234)
        else:
            sigma = np.abs(sigma + 1 / (i + 1) * (mean_alpha - 0.234))
        if trace is not None:
            trace.append(dict(prop=prop.copy(), u_p=np.copy(u_p), x_p=x_p.copy(), ldj_p=ldj_p.copy(),
                              alpha=alpha.copy(), acc=acc.copy(), sigma=float(sigma),
                              mu=None if not tp else mu.copy()))
        cur = track()
        if cur > best:
            cnt, best = 0, cur
        else:
            cnt += 1
            ratio = cap / sigma
            if kind == "rwm_flow":
                ratio = np.minimum(1.0, ratio)
            if cnt >= n_steps * ratio ** 2.0:
                break
        if i >= n_max:
            break
    return dict(u=u, x=x, logdetj=logdetj, logl=logl, logp=logp, efficiency=sigma,
                accept=mean_alpha, steps=i, calls=n_calls, proposal_scale=sigma)


# --------------------------------------------------------------------------------------
# Persistent-sampling weights, ESS, trimming, resampling -- particles.py / tools.py / sampler.py
# --------------------------------------------------------------------------------------
def ps_log_denominator(logl, beta, logz):
    """Running logaddexp over iterations i of beta_i*logl - logz_i (particles.py:221-223).

    Streaming form of ``np.logaddexp.reduce(b, axis=0)``; O(T*N) memory, same summation order.
    Returns the UN-normalised denominator (without the -log T).
    """
    logl = np.asarray(logl, dtype=np.float64)
    acc = logl * beta[0] - logz[0]
    for i in range(1, len(beta)):
        acc = np.logaddexp(acc, logl * beta[i] - logz[i])
    return acc


def ps_logw(logl, beta, logz, beta_final=1.0, normalize=True):
    """Particles.compute_logw_and_logz (particles.py:215-231)."""
    logl = np.asarray(logl, dtype=np.float64)
    beta = np.asarray(beta, dtype=np.float64); logz = np.asarray(logz, dtype=np.float64)
    B = ps_log_denominator(logl, beta, logz) - np.log(len(beta))
    logw = (logl * beta_final - B).reshape(-1)
    logz_new = np.logaddexp.reduce(logw) - np.log(len(logw))
    if normalize:
        logw = logw - np.logaddexp.reduce(logw)
    return logw, logz_new


def ess(weights):
    """tools.py:56-71 (without the in-place normalisation side effect)."""
    w = weights / np.sum(weights)
    return 1.0 / np.sum(w ** 2.0)


def uss(weights, k=None):
    """tools.py:74-93."""
    k = len(weights) if k is None else k
    w = weights / np.sum(weights)
    return np.sum(1.0 - (1.0 - w) ** k)


def compute_ess(logw):
    """tools.py:96-114."""
    e = np.exp(logw - np.max(logw))
    w = e / np.sum(e)
    return 1.0 / np.sum(w * w) / len(w)


def increment_logz(logw):
    """tools.py:117-133."""
    m = np.max(logw)
    return m + np.logaddexp.reduce(logw - m)


def trim_weights(samples, weights, ess_frac=0.99, bins=1000):
    """tools.py:10-53: walk the percentile grid linspace(0,99,bins) downward from 99."""
    w = weights / np.sum(weights)
    total = 1.0 / np.sum(w ** 2.0)
    grid = np.linspace(0, 99, bins)
    i = bins - 1
    while True:
        thr = np.percentile(w, grid[i])
        keep = w >= thr
        wt = w[keep]
        wt = wt / np.sum(wt)
        if (1.0 / np.sum(wt ** 2.0)) / total >= ess_frac:
            break
        i -= 1
    return samples[keep], wt


def trim_weights_sorted(weights, ess_frac=0.99, bins=1000):
    """O(n log n) restatement of trim_weights used to cross-check the device algorithm:
    sort once, suffix sums of w and w^2, evaluate every grid point, take the first (from the
    top) that satisfies the ESS-ratio test.  Returns (keep mask, trimmed weights, grid index)."""
    w = weights / np.sum(weights)
    n = len(w)
    total = 1.0 / np.sum(w ** 2.0)
    srt = np.sort(w)
    s1 = np.cumsum(srt[::-1])[::-1]
    s2 = np.cumsum((srt ** 2.0)[::-1])[::-1]
    grid = np.linspace(0, 99, bins)
    for i in range(bins - 1, -1, -1):
        thr = np.percentile(w, grid[i])
        first = np.searchsorted(srt, thr, side="left")      # first sorted index with w >= thr
        e = s1[first] ** 2 / s2[first]
        if e / total >= ess_frac:
            keep = w >= thr
            wt = w[keep]
            return keep, wt / np.sum(wt), i
    raise RuntimeError("unreachable: the 0th percentile keeps everything")


def systematic_resample(size, weights, u0):
    """tools.py:136-186 with the single uniform ``u0`` = np.random.random() passed in."""
    weights = np.asarray(weights, dtype=np.float64)
    if abs(np.sum(weights) - 1.0) > SQRTEPS:
        weights = weights / np.sum(weights)
    positions = (u0 + np.arange(size)) / size
    idx = np.empty(size, dtype=np.int64)
    j, c = 0, weights[0]
    for i in range(size):
        while positions[i] > c:
            j += 1
            c += weights[j]
        idx[i] = j
    return idx


def multinomial_resample(weights, r):
    """np.random.choice(M, n, True, p=w) (sampler.py:702-703) given its uniforms ``r``:
    cdf = cumsum(p); cdf /= cdf[-1]; searchsorted(cdf, r, 'right')."""
    cdf = np.cumsum(np.asarray(weights, dtype=np.float64))
    cdf /= cdf[-1]
    return cdf.searchsorted(r, side="right").astype(np.int64)


def reweight_select_beta(logl, beta, logz, n_effective, metric="ess"):
    """The beta-selection control flow of Sampler._reweight (sampler.py:739-781), array maths via
    ps_logw.  Returns (beta, logz, ess_est, normalised weights, number of logw probes)."""
    probes = [0]

    def weights_and_ess(b):
        probes[0] += 1
        lw, _ = ps_logw(logl, beta, logz, b)
        w = np.exp(lw - np.max(lw))
        return w, (ess(w) if metric == "ess" else uss(w))

    b_prev = float(beta[-1])
    b_lo, b_hi = b_prev, 1.0
    w_prev, e_prev = weights_and_ess(b_prev)
    w_max, e_max = weights_and_ess(1.0)
    if e_prev <= n_effective:
        b, lz, e = b_prev, float(logz[-1]), e_prev
    elif e_max >= n_effective:
        b, e = 1.0, e_max
        lz = ps_logw(logl, beta, logz, b)[1]
    else:
        while True:
            b = (b_hi + b_lo) * 0.5
            _, e = weights_and_ess(b)
            if abs(e - n_effective) < 0.01 * n_effective or b == 1.0:
                lz = ps_logw(logl, beta, logz, b)[1]
                break
            elif e < n_effective:
                b_hi = b
            else:
                b_lo = b
    lw, _ = ps_logw(logl, beta, logz, b)
    w = np.exp(lw - np.max(lw))
    w /= np.sum(w)
    return b, lz, e, w, probes[0]


def flow_is_evidence(logl, logp, logdetj, logq, boot_idx=None):
    """Importance-sampling evidence + bootstrap error (sampler.py:907-913).
    ``boot_idx`` [B, n] are the bootstrap index draws (np.random.choice(n, n) per row)."""
    logw = logl + logp + logdetj - logq
    n = len(logw)
    logz = np.logaddexp.reduce(logw) - np.log(n)
    if boot_idx is None:
        return logz, None
    boots = np.array([np.logaddexp.reduce(logw[ix]) - np.log(n) for ix in boot_idx])
    return logz, np.std(boots)


# --------------------------------------------------------------------------------------
# Proposal geometry -- geometry.py / student.py  ("next" row f1; host maths)
# --------------------------------------------------------------------------------------
def fit_mvstud(data, tolerance=1e-6, max_iter=100):
    """EM fit of a multivariate Student-t (student.py:5-85). Returns (mu, Sigma, nu)."""
    from scipy import optimize, special
    X = np.asarray(data, dtype=np.float64).T
    dim, n = X.shape
    mu = np.median(X, axis=1)[:, None]
    Sigma = np.cov(X) * (n - 1) / n + (1 / n) * np.diag(np.var(X, axis=1))
    nu, last_nu, it = 20, 0, 0
    while abs(last_nu - nu) > tolerance and it < max_iter:
        it += 1
        diffs = X - mu
        delta = np.sum(diffs * np.linalg.solve(Sigma, diffs), 0)

        def f(v):
            w = (v + dim) / (v + delta)
            return (-special.psi(v / 2) + np.log(v / 2) + np.sum(np.log(w)) / n - np.sum(w) / n + 1
                    + special.psi((v + dim) / 2) - np.log((v + dim) / 2))

        last_nu = nu
        nu = np.inf if f(1e300) >= 0 else optimize.bisect(f, 1e-300, 1e300)
        if nu == np.inf:
            return mu.T[0], Sigma, nu
        w = (nu + dim) / (nu + delta)
        Sigma = np.dot(w * diffs, diffs.T) / n
        mu = (np.sum(w * X, 1) / np.sum(w))[:, None]
    return mu.T[0], Sigma, nu


def geometry_fit(theta, weights, u0):
    """Geometry.fit (geometry.py:31-59) with the systematic-resample uniform ``u0`` explicit."""
    theta = np.asarray(theta)
    out = {}
    if weights is None:
        out["normal_mean"] = np.mean(theta, axis=0)
        out["normal_cov"] = np.cov(theta.T)
        sel = theta
    else:
        out["normal_mean"] = np.average(theta, axis=0, weights=weights)
        out["normal_cov"] = np.cov(theta.T, aweights=weights)
        sel = theta[systematic_resample(len(theta), weights, u0)]
    mu, S, nu = fit_mvstud(sel)
    out["t_mean"], out["t_cov"], out["t_nu"] = mu, S, (nu if np.isfinite(nu) else 1e6)
    return out
