"""Particle history + persistent-sampling importance weights: the reference's
``pocomc.particles.Particles`` (pocomc/particles.py) with the weight maths on the GPU.

The host keeps the reference's append-only dict of per-iteration arrays (API, pickling).  A device
mirror holds ``logl [T, N]`` and the running log-denominator ``den [T, N] =
logaddexp_i(beta_i logl - logz_i)``; appending an iteration folds one term into every row
(O(T N)) instead of rebuilding the reference's [T, T, N] tensor (particles.py:222) per probe, and
each beta probe is one streaming reduction (pmc_ps_reduce).
"""
from __future__ import annotations

import math

import numpy as np
import torch

from . import _lib

KEYS = ("u", "x", "logdetj", "logl", "logp", "logw", "blobs", "iter", "logz", "calls", "steps",
        "efficiency", "ess", "accept", "beta")


class Particles:
    """
    Class to store the particles and their associated weights.

    Parameters
    ----------
    n_particles : int
        Number of particles.
    n_dim : int
        Dimension of the parameter space.
    """

    def __init__(self, n_particles, n_dim):
        self.n_particles = n_particles
        self.n_dim = n_dim
        self.past = {k: [] for k in KEYS}
        self.results_dict = None
        self._reset_device()

    # -- host bookkeeping (particles.py:93-213) ------------------------------------------------
    def _reset_device(self):
        self._d_logl = self._d_den = self._d_beta = self._d_logz = None
        self._t_done = 0

    def __getstate__(self):
        st = self.__dict__.copy()
        for k in ("_d_logl", "_d_den", "_d_beta", "_d_logz"):
            st[k] = None
        st["_t_done"] = 0
        return st

    def update(self, data):
        """Append one iteration (only the reference's keys are stored)."""
        for key in data.keys():
            if key in self.past.keys():
                self.past.get(key).append(data.get(key))

    def pop(self, key):
        _ = self.past.get(key).pop()
        if key in ("logl", "beta", "logz"):
            self._reset_device()

    def get(self, key, index=None, flat=False):
        if index is None:
            if flat:
                return np.concatenate(self.past.get(key))
            return np.asarray(self.past.get(key))
        return self.past.get(key)[index]

    def take_flat(self, key, idx):
        """``self.get(key, flat=True)[idx]`` without materialising the concatenated history: rows are copied
        iteration by iteration, so the cost follows the number of selected rows instead of T x N (the trimmed
        gather of ``Sampler._reweight`` runs once per temperature step and T grows with every step)."""
        arrs = self.past.get(key)
        idx = np.asarray(idx, dtype=np.int64)
        if len(arrs) == 0 or idx.ndim != 1 or (idx.size > 1 and np.any(idx[1:] < idx[:-1])):
            return self.get(key, flat=True)[idx]
        first = np.asarray(arrs[0])
        starts = np.concatenate([[0], np.cumsum([len(a) for a in arrs])])
        if idx.size and (idx[0] < 0 or idx[-1] >= starts[-1]):
            raise IndexError("index out of range of the flattened history")
        dtypes = {np.asarray(a).dtype for a in arrs}
        if len(dtypes) != 1 or any(np.asarray(a).shape[1:] != first.shape[1:] for a in arrs):
            return self.get(key, flat=True)[idx]                      # mixed dtypes / shapes: let numpy decide
        out = np.empty((idx.size,) + first.shape[1:], dtype=first.dtype)
        cut = np.searchsorted(idx, starts)
        for t, a in enumerate(arrs):
            lo, hi = cut[t], cut[t + 1]
            if hi > lo:
                out[lo:hi] = np.asarray(a)[idx[lo:hi] - starts[t]]
        return out

    # -- device mirror --------------------------------------------------------------------------
    def _sync_device(self):
        """Bring logl / den on the GPU up to date with the host history."""
        dev = _lib.device()
        T = len(self.past["logl"])
        if T == 0:
            raise ValueError("no particles stored yet")
        if not (len(self.past["beta"]) == T and len(self.past["logz"]) == T):
            raise ValueError("logl, beta and logz histories must have the same length")
        n = int(np.asarray(self.past["logl"][0]).shape[0])
        if self._d_logl is not None and (self._d_logl.device != dev or self._d_logl.shape[1] != n):
            self._reset_device()
        if self._t_done == T:
            return T, n
        cap = 0 if self._d_logl is None else self._d_logl.shape[0]
        if cap < T:
            new_cap = max(T, 2 * cap, 16)
            logl = torch.empty((new_cap, n), dtype=torch.float64, device=dev)
            den = torch.empty((new_cap, n), dtype=torch.float64, device=dev)
            if self._t_done:
                logl[:self._t_done].copy_(self._d_logl[:self._t_done])
                den[:self._t_done].copy_(self._d_den[:self._t_done])
            self._d_logl, self._d_den = logl, den
        rows = np.ascontiguousarray(np.stack([np.asarray(a, dtype=np.float64) for a in self.past["logl"][self._t_done:T]]))
        self._d_logl[self._t_done:T].copy_(torch.from_numpy(rows))
        self._d_beta = torch.as_tensor(np.asarray(self.past["beta"], dtype=np.float64)).to(dev)
        self._d_logz = torch.as_tensor(np.asarray(self.past["logz"], dtype=np.float64)).to(dev)
        _lib.call("pmc_ps_append", _lib.ptr(self._d_logl), _lib.ptr(self._d_den), _lib.ptr(self._d_beta),
                  _lib.ptr(self._d_logz), int(self._t_done), int(T), n)
        self._t_done = T
        return T, n

    def probe(self, beta_final, uss_k=0):
        """(max logw, sum e, sum e^2, uss sum, T*N) for one beta -- the array maths of
        ``get_weights_and_ess`` (sampler.py:739-746).  Returns host floats + the device stats."""
        T, n = self._sync_device()
        m = T * n
        dev = self._d_logl.device
        scratch = torch.empty(int(_lib.load().pmc_ps_scratch_size(m)), dtype=torch.float64, device=dev)
        out4 = torch.empty(4, dtype=torch.float64, device=dev)
        _lib.call("pmc_ps_reduce", _lib.ptr(self._d_logl), _lib.ptr(self._d_den), float(beta_final), int(T), n,
                  int(uss_k), _lib.ptr(scratch), _lib.ptr(out4))
        h = out4.cpu().numpy()
        return dict(max=float(h[0]), s1=float(h[1]), s2=float(h[2]), uss=float(h[3]), m=m, stats=out4,
                    ess=float(h[1] * h[1] / h[2]), logz=float(h[0] + math.log(h[1]) - math.log(m)))

    def weights_device(self, beta_final, stats=None, want_logw=False):
        """Normalised weights (and optionally normalised logw) over the flattened history, on device."""
        T, n = self._sync_device()
        if stats is None:
            stats = self.probe(beta_final)["stats"]
        dev = self._d_logl.device
        w = torch.empty(T * n, dtype=torch.float64, device=dev)
        lw = torch.empty(T * n, dtype=torch.float64, device=dev) if want_logw else None
        _lib.call("pmc_ps_weights", _lib.ptr(self._d_logl), _lib.ptr(self._d_den), float(beta_final), int(T), n,
                  _lib.ptr(stats), _lib.ptr(w), _lib.ptr(lw))
        return (w, lw) if want_logw else w

    # -- the two questions Sampler._reweight asks beyond the probe; a sharded store (pocomc_b200.sharded) answers
    #    them with an exchange, this one directly
    def weights_global(self, beta_final, stats=None):
        """normalised weights of the WHOLE flattened history ``[T * N]`` on the device"""
        return self.weights_device(beta_final, stats=stats)

    def take_rows(self, key, idx):
        """rows ``idx`` (ascending flat indices into the whole history) of ``key``"""
        return self.take_flat(key, idx)

    def compute_logw_and_logz(self, beta_final=1.0, normalize=True):
        """Persistent-sampling log-weights of every stored particle and the evidence estimate for
        ``beta_final`` (particles.py:215-231):
        logw = beta_f logl - (LSE_i(beta_i logl - logz_i) - log T),  logz = LSE(logw) - log(T N)."""
        p = self.probe(beta_final)
        T, n = self._t_done, self._d_logl.shape[1]
        dev = self._d_logl.device
        stats = p["stats"] if normalize else torch.tensor([0.0, 1.0, 1.0, 0.0], dtype=torch.float64, device=dev)
        lw = torch.empty(T * n, dtype=torch.float64, device=dev)
        _lib.call("pmc_ps_weights", _lib.ptr(self._d_logl), _lib.ptr(self._d_den), float(beta_final), int(T), n,
                  _lib.ptr(stats), None, _lib.ptr(lw))
        return lw.cpu().numpy(), p["logz"]

    def compute_results(self):
        """Stack the history into arrays and attach the final log-weights (particles.py:233-301)."""
        if self.results_dict is None:
            self.results_dict = dict()
            for key in self.past.keys():
                self.results_dict[key] = self.get(key)
            logw, _ = self.compute_logw_and_logz(1.0)
            self.results_dict["logw"] = logw
        return self.results_dict
