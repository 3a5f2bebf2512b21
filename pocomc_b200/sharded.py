"""History sharded by particle -- opt-in (``config.shard_history``); validated on B200 (profiles/r2a_experimental_tests.log).

``Sampler._mutate_sharded`` already splits the mutation step over the ranks but every rank keeps the whole
particle history (T iterations x N particles x (2 D + 3) f64), which does not fit a host at BASELINE configs 4-5.
``ShardedParticles`` stores only this rank's block of particles of every iteration and answers the questions
``Sampler._reweight`` asks with three small exchanges (pocomc_b200.dist):

* beta probe (sampler.py:739-746): local ``pmc_ps_reduce`` -> 24-byte rank-ordered merge -> the same ESS / logZ
  on every rank, hence the same branch of the bisection;
* weights (sampler.py:780-781): local ``pmc_ps_weights`` normalised with the global statistics; the scalar weights
  are all-gathered to ``[T, N]`` (8 bytes per history element) so that the bit-exact trimming of
  ``tools.trim_weights`` can run redundantly on every rank;
* trimmed / resampled rows: global flat indices -> (owner, local row); every rank contributes the rows it owns
  and one ragged all-gather returns them in global index order.

Exchange logic is covered on the CPU over gloo (tests/test_sharded_particles_gloo.py, device kernels replaced by
numpy stand-ins); it has not run on GPUs yet."""
from __future__ import annotations

import math

import numpy as np
import torch

from . import _lib, dist
from .particles import Particles

__all__ = ["ShardedParticles"]


class ShardedParticles(Particles):
    """``Particles`` of one rank: rows ``[lo_r, hi_r)`` of every stored iteration.

    Parameters
    ----------
    n_particles : int
        Number of active particles of the whole run (N).
    n_dim : int
        Dimension of the parameter space.
    counts : sequence of int
        Particles owned by every rank (``dist.shard_counts``); ``update`` expects arrays with ``counts[rank]`` rows.
    rank : int
        This process's rank.
    """

    def __init__(self, n_particles, n_dim, counts, rank):
        super().__init__(n_particles, n_dim)
        self.counts = [int(c) for c in counts]
        self.rank = int(rank)
        if sum(self.counts) != int(n_particles):
            raise ValueError("shard counts do not add up to n_particles")

    PER_PARTICLE = ("u", "x", "logdetj", "logl", "logp", "logw", "blobs")

    def _block(self):
        lo = sum(self.counts[:self.rank])
        return lo, lo + self.counts[self.rank]

    # -- storing and reading back ------------------------------------------------------------------
    def update(self, data):
        """Append one iteration.  Per-particle arrays may arrive whole (N rows: what ``Sampler`` hands over after
        the per-temperature all-gather of the mutated rows) or already cut to this rank's block."""
        lo, hi = self._block()
        cut = {}
        for key, val in data.items():
            if key in self.PER_PARTICLE and val is not None and np.ndim(val) >= 1 and len(val) == self.n_particles \
                    and self.n_particles != hi - lo:
                val = np.asarray(val)[lo:hi]
            cut[key] = val
        super().update(cut)

    def _gather_history(self, key):
        """this rank's ``[T, n_r, ...]`` of ``key`` -> the whole ``[T, N, ...]`` on every rank"""
        local = np.asarray(self.past.get(key))
        if not dist.is_active() or local.ndim < 2:
            return local
        T, n_local = local.shape[:2]
        tail = local.shape[2:]
        rows = np.ascontiguousarray(np.moveaxis(local, 1, 0).reshape(n_local, -1))      # [n_r, T * prod(tail)]
        full = dist.gather_blocks(torch.from_numpy(rows), self.counts).numpy()
        return np.moveaxis(full.reshape((self.n_particles, T) + tail), 0, 1)

    def get(self, key, index=None, flat=False):
        """Like ``Particles.get``; reading the WHOLE history of a per-particle key (``index=None``) is a collective
        that reassembles it in global particle order (posterior(), results)."""
        if index is None and key in self.PER_PARTICLE and len(self.past.get(key)) and self.past.get(key)[0] is not None:
            full = self._gather_history(key)
            return full.reshape((-1,) + full.shape[2:]) if flat else full
        return super().get(key, index=index, flat=flat)

    def compute_logw_and_logz(self, beta_final=1.0, normalize=True):
        p = self.probe(beta_final)
        T, n_local = self._t_done, self._d_logl.shape[1]
        dev = self._d_logl.device
        stats = p["stats"] if normalize else torch.tensor([0.0, 1.0, 1.0, 0.0], dtype=torch.float64, device=dev)
        lw = torch.empty(T * n_local, dtype=torch.float64, device=dev)
        _lib.call("pmc_ps_weights", _lib.ptr(self._d_logl), _lib.ptr(self._d_den), float(beta_final), int(T), n_local,
                  _lib.ptr(stats), None, _lib.ptr(lw))
        return self.global_scalars(lw).cpu().numpy(), p["logz"]

    # -- beta probe ------------------------------------------------------------------------------
    def probe(self, beta_final, uss_k=0):
        if uss_k:
            raise NotImplementedError("the unique-sample-size metric needs a second pass with the global statistics; "
                                      "sharded histories support metric='ess'")
        T, n_local = self._sync_device()
        m = T * self.n_particles
        dev = self._d_logl.device
        scratch = torch.empty(int(_lib.load().pmc_ps_scratch_size(T * n_local)), dtype=torch.float64, device=dev)
        out4 = torch.zeros(4, dtype=torch.float64, device=dev)
        _lib.call("pmc_ps_reduce", _lib.ptr(self._d_logl), _lib.ptr(self._d_den), float(beta_final), int(T), n_local,
                  0, _lib.ptr(scratch), _lib.ptr(out4))
        merged = dist.combine_weight_stats(out4[:3].clone())
        stats = torch.zeros(4, dtype=torch.float64, device=dev)
        stats[:3] = merged.to(dev)
        h = stats.cpu().numpy()
        return dict(max=float(h[0]), s1=float(h[1]), s2=float(h[2]), uss=float("nan"), m=m, stats=stats,
                    ess=float(h[1] * h[1] / h[2]), logz=float(h[0] + math.log(h[1]) - math.log(m)))

    # -- scalars of the whole history --------------------------------------------------------------
    def global_scalars(self, local_flat: torch.Tensor) -> torch.Tensor:
        """``local_flat [T * n_r]`` (iteration-major, what ``weights_device`` returns) -> ``[T * N]`` in the flat
        order of an unsharded history, identical on every rank."""
        T = self._t_done
        n_local = self.counts[self.rank]
        return dist.gather_history_scalars(local_flat.view(T, n_local), self.counts).reshape(-1)

    # -- rows selected by global flat index ----------------------------------------------------------
    def take_flat_global(self, key, idx):
        """``Particles.get(key, flat=True)[idx]`` of the UNSHARDED history (idx ascending), on every rank."""
        idx = np.asarray(idx, dtype=np.int64)
        owner, local = dist.split_history_index(idx, self.n_particles, self.counts)
        mine = np.asarray(self.take_flat(key, local[owner == self.rank]))
        tail = mine.shape[1:]
        width = int(np.prod(tail)) if tail else 1
        rows = torch.from_numpy(np.ascontiguousarray(mine.reshape(mine.shape[0], width)))
        sel_counts = [int(np.sum(owner == r)) for r in range(len(self.counts))]
        got = dist.gather_blocks(rows, sel_counts).numpy()
        where = np.concatenate([np.nonzero(owner == r)[0] for r in range(len(self.counts))])
        out = np.empty((idx.size, got.shape[1]), dtype=got.dtype)
        out[where] = got
        return out.reshape((idx.size,) + tail)

    # -- the store interface of Sampler._reweight ------------------------------------------------------
    def weights_global(self, beta_final, stats=None):
        return self.global_scalars(self.weights_device(beta_final, stats=stats))

    def take_rows(self, key, idx):
        return self.take_flat_global(key, idx)

