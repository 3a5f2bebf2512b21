"""SMC numerics and small utilities with the names of ``pocomc.tools`` (pocomc/tools.py).

Array maths (ESS/USS, trimming, resampling) runs on the GPU through libpmc_b200; inputs and
outputs are numpy arrays like the reference's.  Pure bookkeeping helpers (ProgressBar,
FunctionWrapper, dtype shims) stay on the host.
"""
from __future__ import annotations

import math
import warnings

import numpy as np
import torch

from . import _lib

SQRTEPS = math.sqrt(float(np.finfo(np.float64).eps))

__all__ = ["trim_weights", "effective_sample_size", "unique_sample_size", "compute_ess", "increment_logz",
           "systematic_resample", "ProgressBar", "FunctionWrapper", "torch_to_numpy", "numpy_to_torch",
           "torch_double_to_float", "flow_numpy_wrapper"]


def _dev():
    _lib.require_cuda()
    return torch.device("cuda", torch.cuda.current_device())


def _to_dev(a, dtype=torch.float64):
    if isinstance(a, torch.Tensor):
        return a.to(_dev(), dtype).contiguous()
    return torch.as_tensor(np.ascontiguousarray(a), dtype=dtype).to(_dev())


# ------------------------------------------------------------------------------------------
# device primitives shared with sampler.py / particles.py (operate on CUDA f64 tensors)
# ------------------------------------------------------------------------------------------
def trim_weights_device(w: torch.Tensor, ess: float = 0.99, bins: int = 1000):
    """tools.py:10-53 on a normalised CUDA weight vector: returns (keep mask, trimmed weights).
    Sort + suffix sums evaluate the whole percentile grid in one pass."""
    m = w.numel()
    ws, _ = torch.sort(w)
    scratch = torch.empty(int(_lib.load().pmc_trim_scratch_size(m)), dtype=torch.float64, device=w.device)
    out3 = torch.empty(3, dtype=torch.float64, device=w.device)
    _lib.call("pmc_trim_threshold", _lib.ptr(ws), m, float(ess), int(bins), _lib.ptr(scratch), _lib.ptr(out3))
    keep = w >= out3[0]
    wt = w[keep]
    return keep, wt / wt.sum()


def cumsum_device(w: torch.Tensor) -> torch.Tensor:
    cdf = torch.empty_like(w)
    _lib.call("pmc_cumsum_f64", _lib.ptr(w), _lib.ptr(cdf), w.numel())
    return cdf


def multinomial_resample_device(w: torch.Tensor, r: torch.Tensor) -> torch.Tensor:
    """np.random.choice(M, n, True, p=w) for pre-drawn uniforms r (sampler.py:702-703)."""
    cdf = cumsum_device(w)
    idx = torch.empty(r.numel(), dtype=torch.int64, device=w.device)
    _lib.call("pmc_resample_multinomial", _lib.ptr(cdf), _lib.ptr(r), _lib.ptr(idx), w.numel(), r.numel())
    return idx


def systematic_resample_device(size: int, w: torch.Tensor, u0: float) -> torch.Tensor:
    if abs(float(w.sum().item()) - 1.0) > SQRTEPS:
        w = w / w.sum()
    cdf = cumsum_device(w)
    idx = torch.empty(int(size), dtype=torch.int64, device=w.device)
    _lib.call("pmc_resample_systematic", _lib.ptr(cdf), float(u0), _lib.ptr(idx), w.numel(), int(size))
    return idx


def gather_rows_device(src: torch.Tensor, idx: torch.Tensor) -> torch.Tensor:
    d = 1 if src.dim() == 1 else src.shape[1]
    out = torch.empty((idx.numel(),) + tuple(src.shape[1:]), dtype=torch.float64, device=src.device)
    _lib.call("pmc_gather_rows_f64", _lib.ptr(src), _lib.ptr(idx), _lib.ptr(out), idx.numel(), int(d))
    return out


def weight_stats_device(w: torch.Tensor, uss_k: int = 0) -> torch.Tensor:
    """[sum w, sum w^2, sum 1-(1-w/sum)^k] as a CUDA f64 tensor (tools.py:56-93)."""
    scratch = torch.empty(int(_lib.load().pmc_ps_scratch_size(w.numel())), dtype=torch.float64, device=w.device)
    out3 = torch.empty(3, dtype=torch.float64, device=w.device)
    _lib.call("pmc_weight_stats", _lib.ptr(w), w.numel(), int(uss_k), _lib.ptr(scratch), _lib.ptr(out3))
    return out3


def lse_device(logw, boot_idx=None, n_boot=0, boot_rows=None, seed=None):
    """logsumexp(logw) - log n and the bootstrap replicates logsumexp(logw[idx_b]) - log n (sampler.py:910-913) on the
    GPU; numpy in / (float, numpy) out.  The resampling indices come from one of
      * ``boot_idx``  : an explicit [B, n] index matrix (tests);
      * ``boot_rows`` : a callable ``rows(k) -> [k, n] int64`` drawing the NEXT k rows from the host stream (the
        reference draws ``np.random.choice(n, n)`` row by row): consumed in chunks of 256 rows, so host and device hold
        O(256 n) indices instead of the O(n^2) matrix;
      * ``seed``      : drawn on the device (Philox), no index matrix at all -- ``config.rng_mode == "device"``."""
    lw = _to_dev(logw)
    n = lw.numel()
    scratch = torch.empty(int(_lib.load().pmc_ps_scratch_size(n)), dtype=torch.float64, device=lw.device)
    out = torch.empty(2, dtype=torch.float64, device=lw.device)
    _lib.call("pmc_lse", _lib.ptr(lw), n, _lib.ptr(scratch), _lib.ptr(out))
    boots = None
    if boot_idx is not None:
        idx = torch.as_tensor(np.ascontiguousarray(boot_idx, dtype=np.int64)).to(lw.device)
        res = torch.empty(idx.shape[0], dtype=torch.float64, device=lw.device)
        _lib.call("pmc_lse_bootstrap", _lib.ptr(lw), _lib.ptr(idx), n, idx.shape[0], _lib.ptr(res))
        boots = res.cpu().numpy()
    elif boot_rows is not None and n_boot > 0:
        res = torch.empty(n_boot, dtype=torch.float64, device=lw.device)
        done = 0
        while done < n_boot:
            k = min(256, n_boot - done)
            idx = torch.as_tensor(np.ascontiguousarray(boot_rows(k), dtype=np.int64)).to(lw.device)
            _lib.call("pmc_lse_bootstrap", _lib.ptr(lw), _lib.ptr(idx), n, k, _lib.ptr(res[done:]))
            done += k
        boots = res.cpu().numpy()
    elif seed is not None and n_boot > 0:
        res = torch.empty(n_boot, dtype=torch.float64, device=lw.device)
        _lib.call("pmc_lse_bootstrap_rng", _lib.ptr(lw), n, n_boot, _lib.C.c_uint64(int(seed)), _lib.ptr(res))
        boots = res.cpu().numpy()
    return float(out[0].item()), boots


# ------------------------------------------------------------------------------------------
# reference-named numpy-facing functions
# ------------------------------------------------------------------------------------------
def trim_weights(samples, weights, ess=0.99, bins=1000):
    """Trim samples and weights to a given effective sample size (tools.py:10-53).
    Like the reference, ``weights`` is normalised in place."""
    weights /= np.sum(weights)
    keep, wt = trim_weights_device(_to_dev(weights), ess, bins)
    keep = keep.cpu().numpy()
    return samples[keep], wt.cpu().numpy()


def effective_sample_size(weights):
    """1 / sum(w_normalised^2) (tools.py:56-71); normalises ``weights`` in place like the reference."""
    weights /= np.sum(weights)
    st = weight_stats_device(_to_dev(weights)).cpu().numpy()
    return float(st[0] * st[0] / st[1])


def unique_sample_size(weights, k=None):
    """sum(1 - (1 - w)^k) (tools.py:74-93); normalises ``weights`` in place like the reference."""
    if k is None:
        k = len(weights)
    weights /= np.sum(weights)
    return float(weight_stats_device(_to_dev(weights), int(k)).cpu().numpy()[2])


def compute_ess(logw: np.ndarray):
    """ESS fraction from log-weights (tools.py:96-114)."""
    lw = _to_dev(logw)
    scratch = torch.empty(int(_lib.load().pmc_ps_scratch_size(lw.numel())), dtype=torch.float64, device=lw.device)
    out = torch.empty(2, dtype=torch.float64, device=lw.device)
    _lib.call("pmc_lse", _lib.ptr(lw), lw.numel(), _lib.ptr(scratch), _lib.ptr(out))
    st = weight_stats_device(torch.exp(lw - out[1])).cpu().numpy()
    return float(st[0] * st[0] / st[1] / lw.numel())


def increment_logz(logw: np.ndarray):
    """logsumexp(logw) (tools.py:117-133)."""
    lw = _to_dev(logw)
    scratch = torch.empty(int(_lib.load().pmc_ps_scratch_size(lw.numel())), dtype=torch.float64, device=lw.device)
    out = torch.empty(2, dtype=torch.float64, device=lw.device)
    _lib.call("pmc_lse", _lib.ptr(lw), lw.numel(), _lib.ptr(scratch), _lib.ptr(out))
    return float(out[0].item()) + math.log(lw.numel())


def systematic_resample(size, weights, random_state=None):
    """Systematic resampling (tools.py:136-186); draws its single uniform from the global
    ``np.random`` stream like the reference."""
    if random_state is not None:
        np.random.seed(random_state)
    weights = np.asarray(weights, dtype=np.float64)
    if abs(np.sum(weights) - 1.) > SQRTEPS:
        weights = weights / np.sum(weights)
    u0 = np.random.random()
    w = _to_dev(weights)
    cdf = cumsum_device(w)
    idx = torch.empty(int(size), dtype=torch.int64, device=w.device)
    _lib.call("pmc_resample_systematic", _lib.ptr(cdf), float(u0), _lib.ptr(idx), w.numel(), int(size))
    return idx.cpu().numpy()


class ProgressBar:
    """tqdm progress bar with a stats dictionary (tools.py:189-224)."""

    def __init__(self, show: bool = True, initial=0):
        from tqdm import tqdm
        self.progress_bar = tqdm(desc='Iter', disable=not show, initial=initial)
        self.info = dict()

    def update_stats(self, info):
        self.info = {**self.info, **info}
        self.progress_bar.set_postfix(ordered_dict=self.info)

    def update_iter(self):
        self.progress_bar.update(1)

    def close(self):
        self.progress_bar.close()


class FunctionWrapper(object):
    """Picklable ``f(x, *args, **kwargs)`` (tools.py:227-260)."""

    def __init__(self, f, args, kwargs):
        self.f = f
        self.args = [] if args is None else args
        self.kwargs = {} if kwargs is None else kwargs

    def __call__(self, x):
        return self.f(x, *self.args, **self.kwargs)


def torch_to_numpy(x: torch.Tensor) -> np.ndarray:
    return x.detach().cpu().numpy()


def numpy_to_torch(x: np.ndarray) -> torch.Tensor:
    return torch.tensor(x, dtype=torch.float32)


def torch_double_to_float(x: torch.Tensor, warn: bool = True):
    """f64 -> f32 with the reference's warning; other dtypes are rejected (tools.py:295-316)."""
    if x.dtype == torch.float64 and warn:
        warnings.warn("Float64 data is currently unsupported, casting to Float32. Output will also have type Float32.")
        return x.float()
    elif x.dtype == torch.float32:
        return x
    else:
        raise ValueError(f"Unsupported datatype for input data: {x.dtype}")


class flow_numpy_wrapper:
    """numpy f64 -> f32 flow -> numpy f32; forward negates the log-det (tools.py:318-349)."""

    def __init__(self, flow):
        self.flow = flow

    @torch.no_grad()
    def forward(self, v):
        theta, logdetj = self.flow.forward(numpy_to_torch(v))
        return torch_to_numpy(theta), -torch_to_numpy(logdetj)

    @torch.no_grad()
    def inverse(self, theta):
        v, logdetj = self.flow.inverse(numpy_to_torch(theta))
        return torch_to_numpy(v), torch_to_numpy(logdetj)
