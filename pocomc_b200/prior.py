"""Product prior over independent 1-D ``scipy.stats`` frozen distributions: the reference's
``pocomc.prior.Prior`` (pocomc/prior.py:3-171).  The prior is a host-side black box by contract
(any object with ``logpdf / rvs / bounds / dim`` is accepted, sampler.py:204-207); this class
keeps that API and additionally describes itself to the device fast path (``device_spec``) when
every factor is a frozen ``norm`` or ``uniform``."""
import numpy as np

__all__ = ["Prior"]


class Prior:
    """
    Parameters
    ----------
    dists : list of scipy.stats frozen distributions
        One distribution per parameter; ``len(dists)`` is the dimension.
    """

    def __init__(self, dists=None):
        self.dists = dists

    def logpdf(self, x):
        """Sum over dimensions of ``dist.logpdf(x[:, i])`` (prior.py:70-100); x is [n, dim]."""
        total = np.zeros(len(x))
        for i, dist in enumerate(self.dists):
            total += dist.logpdf(x[:, i])
        return total

    def rvs(self, size=1):
        """[size, dim] sample: one ``dist.rvs(size)`` per dimension in order (prior.py:102-132)."""
        return np.transpose([dist.rvs(size=size) for dist in self.dists])

    @property
    def bounds(self):
        """[dim, 2] support of every factor (prior.py:134-154)."""
        return np.array([dist.support() for dist in self.dists])

    @property
    def dim(self):
        return len(self.dists)

    def device_spec(self):
        """(kind[D] int32, loc[D], scale[D]) with kind 0 = norm, 1 = uniform when every factor is
        one of those two frozen scipy distributions, else ``None`` (host evaluation)."""
        kind, loc, scale = [], [], []
        for dist in self.dists or []:
            name = getattr(getattr(dist, "dist", None), "name", None)
            if name not in ("norm", "uniform") or getattr(dist, "kwds", None) is None:
                return None
            args, kwds = list(dist.args), dict(dist.kwds)
            lo = kwds.get("loc", args[0] if len(args) > 0 else 0.0)
            sc = kwds.get("scale", args[1] if len(args) > 1 else 1.0)
            if np.ndim(lo) or np.ndim(sc):
                return None
            kind.append(0 if name == "norm" else 1)
            loc.append(float(lo))
            scale.append(float(sc))
        if not kind:
            return None
        return np.asarray(kind, np.int32), np.asarray(loc, np.float64), np.asarray(scale, np.float64)
