"""Reparameterisation x (bounded, physical) <-> u (unbounded, standardised): the reference's
``pocomc.scaler.Reparameterize`` (pocomc/scaler.py) with the array maths on the GPU.

numpy in / numpy out like the reference; ``inverse_device`` is the zero-copy entry the MCMC
kernels use.  Only the diagonal affine branch exists (the reference's Sampler never enables the
Cholesky branch, sampler.py:314-318)."""
from __future__ import annotations

import ctypes as C
from typing import List, Union

import numpy as np
import torch

from . import _lib
from .input_validation import assert_array_float, assert_array_within_interval


class Reparameterize:
    """
    Parameters
    ----------
    n_dim : ``int``
        Dimensionality of sampling problem
    bounds : ``np.ndarray`` or ``list`` or ``None``
        Parameter bounds (``[D,2]`` or a ``(low, high)`` pair; non-finite / NaN = unbounded)
    periodic, reflective : ``list``
        Indices of parameters with periodic / reflective boundary conditions
    transform : ``str``
        ``"probit"`` (default) or ``"logit"`` for two-sided bounded parameters
    scale : ``bool``
        Rescale to zero mean and unit variance (default true)
    diagonal : ``bool``
        Must be true (diagonal affine transformation)
    """

    def __init__(self, n_dim: int, bounds: Union[np.ndarray, list] = None, periodic: List[int] = None,
                 reflective: List[int] = None, transform: str = "probit", scale: bool = True, diagonal: bool = True):
        self.ndim = n_dim
        if bounds is None:
            bounds = np.full((self.ndim, 2), np.inf)
        elif len(bounds) == 2 and not np.shape(bounds) == (2, 2):
            # reference quirk (scaler.py:62): a (low, high) pair is tiled through float32
            bounds = np.tile(np.array(bounds, dtype=np.float32).reshape(2, 1), self.ndim).T
        bounds = np.asarray(bounds)
        assert_array_float(bounds)
        self.low = bounds.T[0]
        self.high = bounds.T[1]
        self.periodic = periodic
        self.reflective = reflective
        if transform not in ["logit", "probit"]:
            raise ValueError("Please provide a valid transformation function (e.g. logit or probit)")
        self.transform = transform
        if not diagonal:
            raise NotImplementedError("only the diagonal affine transformation is implemented "
                                      "(the reference Sampler never uses diagonal=False)")
        self.mu = None
        self.sigma = None
        self.scale = scale
        self.diagonal = diagonal
        self._create_masks()
        self._dev = None

    @classmethod
    def adopt(cls, other):
        """A ``pocomc_b200`` scaler with the fitted state of any object that looks like the reference's
        ``pocomc.scaler.Reparameterize`` (attributes ndim, low, high, periodic, reflective, transform, scale, diagonal,
        mu, sigma) -- what lets the reference's own ``Sampler`` call this package's MCMC kernels unchanged."""
        if isinstance(other, cls):
            return other
        if not getattr(other, "diagonal", True):
            raise NotImplementedError("only the diagonal affine transformation is implemented")
        new = cls(int(other.ndim), bounds=np.stack([np.asarray(other.low, np.float64), np.asarray(other.high, np.float64)], axis=1),
                  periodic=other.periodic, reflective=other.reflective, transform=other.transform, scale=bool(other.scale))
        new.mu = None if other.mu is None else np.asarray(other.mu, dtype=np.float64)
        new.sigma = None if other.sigma is None else np.asarray(other.sigma, dtype=np.float64)
        return new

    # -- masks (scaler.py:459-490) ------------------------------------------------------------
    def _create_masks(self):
        lo, hi = np.isfinite(self.low), np.isfinite(self.high)
        self.mask_none = ~lo & ~hi
        self.mask_left = lo & ~hi
        self.mask_right = ~lo & hi
        self.mask_both = lo & hi

    def __getstate__(self):
        st = self.__dict__.copy()
        st["_dev"] = None
        return st

    # -- device parameter block ----------------------------------------------------------------
    def _params(self, scale: bool):
        _lib.require_cuda()
        dev = torch.device("cuda", torch.cuda.current_device())
        key = (dev.index, None if self.mu is None else self.mu.tobytes(), None if self.sigma is None else self.sigma.tobytes())
        if self._dev is None or self._dev["key"] != key:
            kind = np.where(self.mask_both, 3, np.where(self.mask_left, 1, np.where(self.mask_right, 2, 0))).astype(np.int32)
            t = dict(key=key, kind=torch.from_numpy(kind).to(dev),
                     low=torch.as_tensor(np.asarray(self.low, dtype=np.float64)).to(dev),
                     high=torch.as_tensor(np.asarray(self.high, dtype=np.float64)).to(dev), bc=None, mu=None, sigma=None)
            if self.periodic is not None or self.reflective is not None:
                bc = np.zeros(self.ndim, np.int32)
                for i in (self.periodic or []):
                    bc[i] |= 1
                for i in (self.reflective or []):
                    bc[i] |= 2
                t["bc"] = torch.from_numpy(bc).to(dev)
            if self.mu is not None:
                t["mu"] = torch.as_tensor(np.asarray(self.mu, dtype=np.float64)).to(dev)
                t["sigma"] = torch.as_tensor(np.asarray(self.sigma, dtype=np.float64)).to(dev)
            self._dev = t
        t = self._dev
        use_scale = bool(scale and self.scale)
        if use_scale and t["mu"] is None:
            raise RuntimeError("Reparameterize.fit must be called before forward/inverse")
        sc = _lib.PmcScaler(kind=_lib.ptr(t["kind"]), bc=None, low=_lib.ptr(t["low"]), high=_lib.ptr(t["high"]),
                            mu=_lib.ptr(t["mu"]) if use_scale else None, sigma=_lib.ptr(t["sigma"]) if use_scale else None,
                            log_sigma_sum=float(np.sum(np.log(self.sigma))) if use_scale else 0.0,
                            logit=1 if self.transform == "logit" else 0, scale=1 if use_scale else 0)
        return sc, t, dev

    # -- public API ------------------------------------------------------------------------------
    def apply_boundary_conditions_x(self, x: np.ndarray):
        """Periodic wrap then reflection (scaler.py:84-157)."""
        if self.periodic is None and self.reflective is None:
            return x
        _, t, dev = self._params(False)
        xd = torch.as_tensor(np.ascontiguousarray(x, dtype=np.float64)).to(dev)
        _lib.call("pmc_apply_bc", _lib.ptr(xd), _lib.ptr(t["bc"]), _lib.ptr(t["low"]), _lib.ptr(t["high"]),
                  xd.shape[0], xd.shape[1])
        return xd.cpu().numpy()

    def fit(self, x: np.ndarray):
        """Learn mean and standard deviation of the unbounded variables (scaler.py:159-178)."""
        assert_array_within_interval(x, self.low, self.high)
        v = self._forward(x)
        self.mu = np.mean(v, axis=0)
        self.sigma = np.std(v, axis=0)

    def _forward_device(self, x: torch.Tensor, scale: bool) -> torch.Tensor:
        sc, _, _ = self._params(scale)
        u = torch.empty_like(x)
        _lib.call("pmc_scaler_forward", _lib.ptr(x), C.byref(sc), _lib.ptr(u), x.shape[0], x.shape[1])
        return u

    def _forward(self, x: np.ndarray):
        """Bounded -> unbounded without the affine part (scaler.py:228-247)."""
        _lib.require_cuda()
        xd = torch.as_tensor(np.ascontiguousarray(x, dtype=np.float64)).cuda()
        return self._forward_device(xd, scale=False).cpu().numpy()

    def forward(self, x: np.ndarray, check_input=True):
        """x -> u (scaler.py:180-202)."""
        if check_input:
            assert_array_within_interval(x, self.low, self.high)
        _lib.require_cuda()
        xd = torch.as_tensor(np.ascontiguousarray(x, dtype=np.float64)).cuda()
        return self._forward_device(xd, scale=True).cpu().numpy()

    def inverse_device(self, u: torch.Tensor, with_bc: bool = False):
        """u (CUDA f32 or f64 [N,D]) -> (u_out f64, x f64, logdetj f64 [N], finite u8 [N]) on device.
        ``with_bc`` applies the boundary wrap + re-forward + re-inverse of mcmc.py:94-97."""
        sc, t, dev = self._params(True)
        if with_bc and t["bc"] is not None:
            sc.bc = _lib.ptr(t["bc"])
        n, d = u.shape
        u_out = torch.empty((n, d), dtype=torch.float64, device=dev)
        x = torch.empty((n, d), dtype=torch.float64, device=dev)
        logdetj = torch.empty(n, dtype=torch.float64, device=dev)
        finite = torch.empty(n, dtype=torch.uint8, device=dev)
        _lib.call("pmc_scaler_inverse", 1 if u.dtype == torch.float32 else 0, _lib.ptr(u), C.byref(sc), _lib.ptr(u_out),
                  _lib.ptr(x), _lib.ptr(logdetj), _lib.ptr(finite), n, d)
        return u_out, x, logdetj, finite

    def inverse(self, u: np.ndarray):
        """u -> (x, log|dx/du|) (scaler.py:204-226).  f32 input is promoted like numpy does."""
        _lib.require_cuda()
        u = np.ascontiguousarray(u)
        if u.dtype not in (np.float32, np.float64):
            u = u.astype(np.float64)
        _, x, logdetj, _ = self.inverse_device(torch.from_numpy(u).cuda())
        return x.cpu().numpy(), logdetj.cpu().numpy()
