"""Preconditioned Monte Carlo driver with the reference's ``pocomc.Sampler`` API
(pocomc/sampler.py:154-1061).  The control flow (temperature ladder, bisection, early stopping,
checkpointing, progress statistics) is host Python exactly as in the reference; every array
operation of the hot path -- persistent-sampling weights, ESS/USS, trimming, resampling, the
MCMC kernels, the flow and the evidence reductions -- runs on the GPU through libpmc_b200.
The user's ``likelihood`` stays a host-side black box (numpy in, numpy out)."""
from __future__ import annotations

import os
import warnings
from pathlib import Path
from typing import Union

import dill
import numpy as np
import torch

from . import _lib, config, dist
from .flow import Flow
from .geometry import Geometry
from .mcmc import pcn, preconditioned_pcn, preconditioned_rwm, rwm
from .particles import Particles
from .scaler import Reparameterize
from .tools import (FunctionWrapper, ProgressBar, flow_numpy_wrapper, gather_rows_device, lse_device,
                    multinomial_resample_device, numpy_to_torch, systematic_resample, torch_to_numpy,
                    trim_weights, trim_weights_device, weight_stats_device)

__all__ = ["Sampler"]


def configure_threads(pytorch_threads=None):
    """pocomc/threading.py:3-21."""
    if pytorch_threads:
        torch.set_num_threads(pytorch_threads)


def _uss_of_equal_weights(n: int, k: int) -> float:
    """unique_sample_size(np.ones(n), k) (tools.py:74-93) -- constructor-time scalar, kept on the
    host so a Sampler can be built before a device is selected."""
    w = np.ones(n)
    w /= np.sum(w)
    return float(np.sum(1.0 - (1.0 - w) ** k))


class Sampler:
    r"""Preconditioned Monte Carlo sampler (see ``pocomc.Sampler`` for the meaning of every
    argument; names, defaults and validation errors are the reference's, sampler.py:154-373).

    Device extras (no counterpart in the reference): a ``likelihood`` object exposing
    ``device(x, finite, out)`` and a prior exposing ``device_spec()`` are evaluated on the GPU when
    ``pocomc_b200.config.device_callbacks`` is true, removing the per-step PCIe round trip.
    """

    def __init__(self, prior: callable, likelihood: callable, n_dim: int = None, n_effective: int = 512,
                 n_active: int = 256, likelihood_args: list = None, likelihood_kwargs: dict = None,
                 vectorize: bool = False, blobs_dtype: str = None, periodic: list = None, reflective: list = None,
                 transform: str = "probit", pool=None, pytorch_threads=1, flow='nsf6', train_config: dict = None,
                 train_frequency: int = None, precondition: bool = True, dynamic: bool = True, metric: str = 'ess',
                 n_prior: int = None, sample: str = 'tpcn', n_steps: int = None, n_max_steps: int = None,
                 resample: str = 'mult', output_dir: str = None, output_label: str = None, random_state: int = None,
                 n_ess: int = None):
        if n_ess is not None:
            n_effective = n_ess
            warnings.warn("n_ess is deprecated. Use n_effective instead.", DeprecationWarning, stacklevel=2)
        if random_state is None and dist.is_active():
            # particle-sharded runs are REPLICATED state machines: every rank must draw the same prior samples, resampling
            # indices and MCMC noise, or the ranks take different branches (and collectives deadlock).  Without a seed from
            # the user, rank 0 picks one for everybody.
            random_state = int(dist.broadcast_from_rank0(np.array([np.random.randint(0, 2 ** 31 - 1)], dtype=np.int64))[0])
        if random_state is not None:                                         # sampler.py:195-197
            np.random.seed(random_state)
            torch.manual_seed(random_state)
        self.random_state = random_state
        configure_threads(pytorch_threads=pytorch_threads)

        self.prior = prior
        self.log_prior = self.prior.logpdf
        self.sample_prior = self.prior.rvs
        self.bounds = self.prior.bounds
        self.log_likelihood = FunctionWrapper(likelihood, likelihood_args, likelihood_kwargs)
        self.blobs_dtype = blobs_dtype
        self.have_blobs = blobs_dtype is not None
        self.n_dim = self.prior.dim if n_dim is None else int(n_dim)

        if n_active is None and n_effective is None:
            raise ValueError("At least one of n_active or n_effective must be provided.")
        self.n_active = int(n_effective / 2) if n_active is None else int(n_active)
        self.n_effective = int(2 * n_active) if n_effective is None else int(n_effective)
        self.n_steps = int(self.n_dim // 2) if n_steps is None else int(n_steps)
        self.n_max_steps = 10 * self.n_steps if n_max_steps is None else int(n_max_steps)
        self.n_total = None
        self.n_evidence = None
        self.particles = Particles(n_active, n_dim)
        if config.shard_history and dist.is_active() and self.n_active >= dist.world()[1]:
            # experimental: every rank keeps only its block of every stored iteration; same
            # block boundaries as _mutate_sharded
            from .sharded import ShardedParticles
            rank, ws = dist.world()
            align = 256 if self.n_active // 256 >= ws else 1
            self.particles = ShardedParticles(self.n_active, self.n_dim, dist.shard_counts(self.n_active, ws, align), rank)
        self.t = 0

        self.pool = pool
        if pool is None:
            self.distribute = map
        elif isinstance(pool, int) and pool > 1:
            from multiprocess import Pool
            self.pool = Pool(pool)
            self.distribute = self.pool.map
        else:
            self.distribute = pool.map
        self.vectorize = vectorize
        if self.vectorize and self.have_blobs:
            raise ValueError("Cannot vectorize likelihood with blobs.")

        self.u_geometry = Geometry()
        self.theta_geometry = Geometry()
        self.flow = Flow(self.n_dim, flow)
        self.train_config = dict(validation_split=0.5, epochs=5000, batch_size=np.minimum(self.n_effective // 2, 512),
                                 patience=int(self.n_dim), learning_rate=1e-3, annealing=False, gaussian_scale=None,
                                 laplace_scale=None, noise=None, shuffle=True, clip_grad_norm=1.0, verbose=0)
        if train_config is not None:
            self.train_config.update(train_config)
        if train_frequency is None:
            self.train_frequency = np.maximum(self.n_effective // (self.n_active * 2), 1)
        else:
            self.train_frequency = int(train_frequency)
        self.flow_untrained = True

        if transform not in ['probit', 'logit']:
            raise ValueError(f"Invalid transform {transform}. Options are 'probit' or 'logit'.")
        self.scaler = Reparameterize(self.n_dim, bounds=self.bounds, periodic=periodic, reflective=reflective,
                                     transform=transform)
        self.output_dir = Path("states") if output_dir is None else output_dir
        self.output_label = "pmc" if output_label is None else output_label
        self.preconditioned = precondition
        if metric not in ['ess', 'uss']:
            raise ValueError(f"Invalid metric {metric}. Options are 'ess' or 'uss'.")
        self.metric = metric
        self.dynamic = dynamic
        self.dynamic_ratio = _uss_of_equal_weights(self.n_effective, self.n_active) / self.n_active
        if sample not in ['tpcn', 'rwm']:
            raise ValueError(f"Invalid sample {sample}. Options are 'tpcn' or 'rwm'.")
        self.sample = sample
        self.proposal_scale = 2.38 / self.n_dim ** 0.5
        if resample not in ['mult', 'syst']:
            raise ValueError(f"Invalid resample {resample}. Options are 'mult' or 'syst'.")
        self.resample = resample
        if n_prior is None:
            self.n_prior = int(2 * np.maximum(self.n_effective // self.n_active, 1) * self.n_active)
        else:
            self.n_prior = int(np.maximum(n_prior / self.n_active, 1) * self.n_active)
        self.prior_samples = None
        self.logz = None
        self.logz_err = None
        self.current_particles = None
        self.warmup = True
        self.calls = 0
        self.progress = None
        self.pbar = None

    # ------------------------------------------------------------------------------------------
    # run loop (sampler.py:375-524)
    # ------------------------------------------------------------------------------------------
    def _history_stats(self):
        last = lambda k: self.particles.get(k, -1)
        return dict(beta=last("beta"), calls=last("calls"), ESS=last("ess"), logZ=last("logz"),
                    logP=np.mean(last("logp") + last("logl")), acc=last("accept"), steps=last("steps"),
                    eff=last("efficiency"))

    def _maybe_save(self, save_every, t0):
        if save_every is not None and (self.t - t0) % int(save_every) == 0 and self.t != t0:
            self.save_state(Path(self.output_dir) / f'{self.output_label}_{self.t}.state')

    def run(self, n_total: int = 4096, n_evidence: int = 4096, progress: bool = True,
            resume_state_path: Union[str, Path] = None, save_every: int = None):
        r"""Run Preconditioned Monte Carlo until ``n_total`` effectively independent samples at
        beta = 1 have been collected; see ``pocomc.Sampler.run``."""
        if resume_state_path is not None:
            self.load_state(resume_state_path)
            t0 = self.t
            self.pbar = ProgressBar(self.progress, initial=t0)
            self.pbar.update_stats(self._history_stats())
        else:
            t0 = self.t
            self.progress = progress and dist.world()[0] == 0          # sharded runs: rank 0 owns the progress bar
            self.pbar = ProgressBar(self.progress)
            self.pbar.update_stats(dict(beta=0.0, calls=self.calls, ESS=self.n_effective, logZ=0.0, logP=0.0,
                                        acc=0.0, steps=0, eff=0.0))
        self.n_total = int(n_total)
        self.n_evidence = int(n_evidence)

        if self.prior_samples is None:
            self.prior_samples = self.sample_prior(self.n_prior)
            if dist.is_active():             # a user prior with its own RNG must not split the replicas
                self.prior_samples = dist.broadcast_from_rank0(np.asarray(self.prior_samples, dtype=np.float64))
            self.scaler.fit(self.prior_samples)

        if self.warmup:                                                    # sampler.py:442-489
            for i in range(self.n_prior // self.n_active):
                self._maybe_save(save_every, t0)
                x = self.prior_samples[i * self.n_active:(i + 1) * self.n_active]
                u = self.scaler.forward(x)
                logdetj = self.scaler.inverse(u)[1]
                logp = self.log_prior(x)
                logl, blobs = self._log_like(x)
                self.calls += self.n_active
                bad = np.isinf(logl)
                if np.any(bad):
                    every = np.arange(len(x))
                    lost, good = every[bad], every[~bad]
                    src = np.random.choice(good, size=len(lost), replace=True)
                    for arr in (x, u, logdetj, logp, logl):
                        arr[lost] = arr[src]
                    if self.have_blobs:
                        blobs[lost] = blobs[src]
                self.current_particles = dict(u=u, x=x, logl=logl, logp=logp, logdetj=logdetj,
                                              logw=-1e300 * np.ones(self.n_active), blobs=blobs, iter=self.t,
                                              calls=self.calls, steps=1, efficiency=1.0, ess=self.n_effective,
                                              accept=1.0, beta=0.0, logz=0.0)
                self.particles.update(self.current_particles)
                stats = self._history_stats()
                stats["ESS"] = int(stats["ESS"])
                self.pbar.update_stats(stats)
                self.pbar.update_iter()
                self.t += 1
            self.warmup = False

        while self._not_termination(self.current_particles):
            self._maybe_save(save_every, t0)
            self.current_particles = self._reweight(self.current_particles)
            self.current_particles = self._train(self.current_particles)
            self.current_particles = self._resample(self.current_particles)
            self.current_particles = self._mutate(self.current_particles)
            self.particles.update(self.current_particles)

        if self.n_evidence > 0 and self.preconditioned:
            self._compute_evidence(self.n_evidence)
        else:
            self.logz = self.particles.probe(1.0)["logz"]
            self.logz_err = None
        if save_every is not None:
            self.save_state(Path(self.output_dir) / f'{self.output_label}_final.state')
        self.pbar.close()

    def _ess_of_probe(self, p):
        return p["ess"] if self.metric == 'ess' else p["uss"]

    def _probe(self, beta):
        """get_weights_and_ess (sampler.py:739-746) as one fused device reduction."""
        m = len(self.particles.past["logl"]) * len(self.particles.past["logl"][0])
        return self.particles.probe(beta, uss_k=m if self.metric == 'uss' else 0)

    def _not_termination(self, current_particles):
        """sampler.py:526-547."""
        ess = self._ess_of_probe(self._probe(1.0))
        return 1.0 - current_particles.get("beta") >= 1e-4 or ess < self.n_total

    # ------------------------------------------------------------------------------------------
    # SMC steps
    # ------------------------------------------------------------------------------------------
    def _mutate(self, current_particles):
        """sampler.py:550-633."""
        blobs = current_particles.get("blobs").copy() if self.have_blobs else None
        state_dict = dict(u=current_particles.get("u").copy(), x=current_particles.get("x").copy(),
                          logdetj=current_particles.get("logdetj").copy(), logp=current_particles.get("logp").copy(),
                          logl=current_particles.get("logl").copy(), beta=current_particles.get("beta"), blobs=blobs)
        function_dict = dict(loglike=self._log_like, logprior=self.log_prior, scaler=self.scaler, flow=self.flow,
                             u_geometry=self.u_geometry, theta_geometry=self.theta_geometry)
        if config.device_callbacks and not self.have_blobs:
            like_dev = getattr(self.log_likelihood.f, "device", None)
            spec = self.prior.device_spec() if hasattr(self.prior, "device_spec") else None
            if like_dev is not None and spec is not None and not self.log_likelihood.args and not self.log_likelihood.kwargs:
                from .synthetic import DevicePrior
                function_dict["loglike_device"] = like_dev
                function_dict["logprior_device"] = DevicePrior(*spec)
        option_dict = dict(n_max=self.n_max_steps, n_steps=self.n_steps, progress_bar=self.pbar,
                           proposal_scale=self.proposal_scale)
        kernel = {(True, "tpcn"): preconditioned_pcn, (True, "rwm"): preconditioned_rwm,
                  (False, "tpcn"): pcn, (False, "rwm"): rwm}[(bool(self.preconditioned), self.sample)]
        if dist.is_active() and self.n_active >= dist.world()[1]:
            results = self._mutate_sharded(kernel, state_dict, function_dict, option_dict)
        else:
            results = kernel(state_dict, function_dict, option_dict)
        for key in ("u", "x", "logdetj", "logl", "logp"):
            current_particles[key] = results.get(key).copy()
        if self.have_blobs:
            current_particles["blobs"] = results.get('blobs').copy()
        current_particles["efficiency"] = results.get('efficiency') / (2.38 / self.n_dim ** 0.5)
        current_particles["steps"] = results.get('steps')
        current_particles["accept"] = results.get('accept')
        current_particles["calls"] = current_particles.get("calls") + results.get('calls')
        self.calls = current_particles.get("calls")
        self.proposal_scale = results.get('proposal_scale')
        return current_particles

    def _mutate_sharded(self, kernel, state_dict, function_dict, option_dict):
        """One process per GPU: every rank holds the same replicated sampler (same
        ``random_state``, same history, same flow) and mutates only its contiguous block of the active
        particles -- so the host likelihood is called on 1/G of the rows per rank.  Per MCMC step the
        ranks exchange the fixed-size block partials behind sigma / mu / stop rule (mcmc.McmcEngine);
        per temperature step ONE rank-ordered all-gather returns the mutated rows to every rank.
        Shard boundaries are multiples of the accept kernel's 256-row blocks whenever every rank can
        get at least one block, which makes the run bit-identical to the single-GPU one (host RNG,
        mean_mode 0)."""
        rank, ws = dist.world()
        n, d = state_dict["x"].shape
        align = 256 if n // 256 >= ws else 1
        lo, hi = dist.shard_range(n, rank, ws, align)
        counts = dist.shard_counts(n, ws, align)
        lib = _lib.load()
        blocks = [int(lib.pmc_mh_partials_size(c, d)) // (d + 4) for c in counts]
        tp = self.sample == "tpcn"
        tot = state_dict["logl"] + state_dict["logp"] + (0.0 if tp else state_dict["logdetj"])
        local = {k: (v[lo:hi] if isinstance(v, np.ndarray) else v) for k, v in state_dict.items()}
        opts = dict(option_dict, shard=(lo, n, blocks), best0=float(np.mean(tot)),
                    progress_bar=option_dict["progress_bar"] if rank == 0 else None)
        res = kernel(local, function_dict, opts)
        dev = torch.device("cuda", torch.cuda.current_device())
        pack = np.concatenate([res["u"], res["x"], res["logdetj"][:, None], res["logl"][:, None], res["logp"][:, None]], axis=1)
        full = dist.gather_blocks(torch.from_numpy(pack).to(dev), counts).cpu().numpy()
        out = dict(res)
        out["u"], out["x"] = np.ascontiguousarray(full[:, :d]), np.ascontiguousarray(full[:, d:2 * d])
        out["logdetj"], out["logl"], out["logp"] = (np.ascontiguousarray(full[:, 2 * d + i]) for i in range(3))
        # host-evaluated likelihoods count this rank's rows; the device-callback path reads the controller block, which the
        # fixed-order reduction already made global
        out["calls"] = res["calls"] if res.get("calls_global") else dist.allreduce_sum_int(res["calls"])
        if self.have_blobs:
            parts = [None] * ws
            torch.distributed.all_gather_object(parts, res["blobs"])
            out["blobs"] = np.concatenate(parts, axis=0)
        return out

    def _train(self, current_particles):
        """sampler.py:636-678."""
        u = current_particles.get("u")
        w = current_particles.get("weights")
        cfg = self.train_config
        if self.preconditioned and (self.t % self.train_frequency == 0 or current_particles.get("beta") == 1.0
                                    or self.flow_untrained):
            self.flow_untrained = False
            self.flow.fit(numpy_to_torch(u), weights=numpy_to_torch(w), validation_split=cfg["validation_split"],
                          epochs=cfg["epochs"], batch_size=int(np.minimum(len(u) // 2, cfg["batch_size"])),
                          gaussian_scale=cfg["gaussian_scale"], laplace_scale=cfg["laplace_scale"],
                          patience=cfg["patience"], learning_rate=cfg["learning_rate"], annealing=cfg["annealing"],
                          noise=cfg["noise"], shuffle=cfg["shuffle"], clip_grad_norm=cfg["clip_grad_norm"],
                          verbose=cfg["verbose"])
            theta = flow_numpy_wrapper(self.flow).forward(u)[0]
            self.theta_geometry.fit(theta, weights=w)
        else:
            self.u_geometry.fit(u, weights=w)
        return current_particles

    def _resample(self, current_particles):
        """sampler.py:680-715: multinomial (np.random.choice(p=w) == searchsorted of the sequential
        f64 cdf with n_active uniforms) or systematic; indices computed on the GPU, bit-exact."""
        weights = current_particles.get("weights")
        if self.resample == 'mult':
            r = np.random.random_sample(self.n_active)
            dev = torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else None
            _lib.require_cuda()
            idx = multinomial_resample_device(torch.as_tensor(np.ascontiguousarray(weights, dtype=np.float64)).to(dev),
                                              torch.from_numpy(r).to(dev)).cpu().numpy()
        else:
            idx = systematic_resample(self.n_active, weights=weights)
        for key in ("u", "x", "logdetj", "logl", "logp"):
            current_particles[key] = current_particles.get(key)[idx]
        if self.have_blobs:
            current_particles["blobs"] = current_particles.get("blobs")[idx]
        return current_particles

    def _reweight(self, current_particles):
        """Choose the next temperature by bisection on the ESS of the persistent-sampling weights,
        adapt n_effective, trim and gather the surviving history (sampler.py:717-805)."""
        self.t += 1
        self.pbar.update_iter()
        beta_prev = self.particles.get("beta", index=-1)
        beta_max, beta_min = 1.0, float(np.copy(beta_prev))

        p_prev = self._probe(beta_prev)
        p_max = self._probe(beta_max)
        ess_prev, ess_max = self._ess_of_probe(p_prev), self._ess_of_probe(p_max)
        if ess_prev <= self.n_effective:
            beta, p_sel, ess_est = beta_prev, p_prev, ess_prev
            logz = self.particles.get("logz", index=-1)
        elif ess_max >= self.n_effective:
            beta, p_sel, ess_est = beta_max, p_max, ess_max
            logz = p_max["logz"]
        else:
            while True:
                beta = (beta_max + beta_min) * 0.5
                p_sel = self._probe(beta)
                ess_est = self._ess_of_probe(p_sel)
                if np.abs(ess_est - self.n_effective) < 0.01 * self.n_effective or beta == 1.0:
                    logz = p_sel["logz"]
                    break
                elif ess_est < self.n_effective:
                    beta_max = beta
                else:
                    beta_min = beta
        self.pbar.update_stats(dict(beta=beta, ESS=int(ess_est), logZ=logz))

        w_dev = self.particles.weights_global(beta, stats=p_sel["stats"])
        if self.dynamic:                                                   # sampler.py:783-790
            n_unique = float(weight_stats_device(w_dev, int(self.n_active)).cpu().numpy()[2])
            if n_unique < self.n_active * (0.95 * self.dynamic_ratio):
                self.n_effective = int(self.n_active / n_unique * self.n_effective)
            elif n_unique > self.n_active * np.minimum(1.05 * self.dynamic_ratio, 1.0):
                self.n_effective = int(n_unique / self.n_active * self.n_effective)
        keep, w_trim = trim_weights_device(w_dev, ess=0.99, bins=1000)
        idx = torch.nonzero(keep).squeeze(1).cpu().numpy()
        for key in ("u", "x", "logdetj", "logl", "logp"):
            current_particles[key] = self.particles.take_rows(key, idx)
        if self.have_blobs:
            current_particles["blobs"] = self.particles.take_rows("blobs", idx)
        current_particles["logz"] = logz
        current_particles["beta"] = beta
        current_particles["weights"] = w_trim.cpu().numpy()
        current_particles["ess"] = ess_est
        return current_particles

    def _log_like(self, x):
        """Host black box (sampler.py:807-861): one vectorised call, or one call per row (through ``pool.map`` when a
        pool was given).  A row result may be ``logl`` or ``(logl, blob, ...)``; blobs come back as one array."""
        if self.vectorize:
            return self.log_likelihood(x), None
        mapper = self.distribute if self.pool is not None else map
        rows = list(mapper(self.log_likelihood, x))
        return _split_blobs(rows, self.blobs_dtype, self)

    # ------------------------------------------------------------------------------------------
    # evidence / posterior / results
    # ------------------------------------------------------------------------------------------
    def evidence(self):
        """(logZ, error) -- sampler.py:863-867."""
        return self.logz, self.logz_err

    def _compute_evidence(self, n=5_000):
        """Importance-sampling evidence with the trained flow as proposal (sampler.py:869-920);
        the log-sum-exp and the max(n,1000)-fold bootstrap run on the GPU."""
        with torch.no_grad():
            theta_q, logq = self.flow.sample(n)
            theta_q = torch_to_numpy(theta_q)
            logq = torch_to_numpy(logq)
        x_q, logdetj = self.scaler.inverse(theta_q)
        logp = self.log_prior(x_q)
        ok = np.isfinite(logp)
        x_q, logdetj, logq, logp = x_q[ok], logdetj[ok], logq[ok], logp[ok]
        logl, _ = self._log_like(x_q)
        logw = logl + logp + logdetj - logq
        m = len(logw)
        n_boot = int(np.maximum(n, 1000))
        if config.rng_mode == "device":
            # resampling indices drawn inside the kernel: no index matrix; one host draw keys the counters
            logz, boots = lse_device(logw, n_boot=n_boot, seed=int(np.random.randint(0, 2 ** 62)))
        else:
            # bit-faithful host stream, reference order (one np.random.choice(m, m) per replicate), 256 rows at a time
            logz, boots = lse_device(logw, n_boot=n_boot, boot_rows=lambda k: np.stack([np.random.choice(m, m) for _ in range(k)]))
        self.calls += m
        self.pbar.update_stats(dict(calls=self.calls))
        self.logz = logz
        self.logz_err = float(np.std(boots))
        return self.logz, self.logz_err

    def __getstate__(self):
        """Pickle everything but the pool (sampler.py:922-939)."""
        state = self.__dict__.copy()
        try:
            if state['pool'] is not None:
                del state['pool']
                del state['distribute']
        except Exception:
            pass
        return state

    def posterior(self, resample=False, return_blobs=False, trim_importance_weights=True, return_logw=False,
                  ess_trim=0.99, bins_trim=1_000):
        """Weighted (or resampled) posterior samples from the whole history (sampler.py:941-1009)."""
        if return_blobs and not self.have_blobs:
            raise ValueError("No blobs available.")
        samples = self.particles.get("x", flat=True)
        logl = self.particles.get("logl", flat=True)
        logp = self.particles.get("logp", flat=True)
        blobs = self.particles.get("blobs", flat=True) if return_blobs else None
        logw, _ = self.particles.compute_logw_and_logz(1.0)
        weights = np.exp(logw)
        if trim_importance_weights:
            idx, weights = trim_weights(np.arange(len(samples)), weights, ess=ess_trim, bins=bins_trim)
            samples, logl, logp, logw = samples[idx], logl[idx], logp[idx], logw[idx]
            if return_blobs:
                blobs = blobs[idx]
        if resample:
            if self.resample == 'mult':
                r = np.random.random_sample(len(samples))
                dev = torch.device("cuda", torch.cuda.current_device())
                pick = multinomial_resample_device(torch.as_tensor(weights).to(dev), torch.from_numpy(r).to(dev)).cpu().numpy()
            else:
                pick = systematic_resample(len(weights), weights=weights)
            out = (samples[pick], logl[pick], logp[pick])
            return out + (blobs[pick],) if return_blobs else out
        out = (samples, logw if return_logw else weights, logl, logp)
        return out + (blobs,) if return_blobs else out

    @property
    def results(self):
        """Stacked history dictionary (sampler.py:1011-1021)."""
        return self.particles.compute_results()

    # ------------------------------------------------------------------------------------------
    # checkpointing (sampler.py:1023-1061)
    # ------------------------------------------------------------------------------------------
    def _state_path(self, path):
        """Replicated runs: one file, written by rank 0.  Particle-sharded history (config.shard_history): every rank owns a
        slice of the history, so rank r > 0 writes / reads ``<path>.rank<r>`` next to rank 0's ``<path>``."""
        rank, ws = dist.world()
        from .sharded import ShardedParticles
        sharded = ws > 1 and isinstance(self.particles, ShardedParticles)
        p = Path(path)
        return rank, ws, sharded, (p if rank == 0 or not sharded else p.with_name(p.name + f".rank{rank}"))

    def save_state(self, path: Union[str, Path]):
        """Atomic dill dump of the sampler's attributes (device mirrors are dropped by the
        members' own ``__getstate__`` and rebuilt lazily after ``load_state``); sampler.py:1023-1049."""
        rank, ws, sharded, mine = self._state_path(path)
        if ws > 1 and not sharded and rank != 0:
            dist.barrier()                        # rank 0 writes the (identical) replicated state
            return
        if rank == 0:
            print(f'Saving PMC state to {path}')
        mine.parent.mkdir(exist_ok=True)
        temp_path = mine.with_name(mine.name + '.temp')
        with open(temp_path, 'wb') as f:
            state = self.__dict__.copy()
            del state['pbar']
            try:
                if state['pool'] is not None:
                    del state['pool']
                    del state['distribute']
            except BaseException as e:
                print(e)
            dill.dump(file=f, obj=state)
            f.flush()
            os.fsync(f.fileno())
        os.rename(temp_path, mine)
        dist.barrier()

    def load_state(self, path: Union[str, Path]):
        """sampler.py:1051-1061; under a sharded history every rank reads its own slice back."""
        _, _, _, mine = self._state_path(path)
        with open(mine if mine.exists() else path, 'rb') as f:
            state = dill.load(file=f)
        self.__dict__ = {**self.__dict__, **state}


def _split_blobs(rows, blobs_dtype, owner=None):
    """Separate per-row likelihood results into (logl [n], blobs | None).

    Scalars (anything without ``len``) mean "no blobs".  With tuple-like rows the first entry is the
    log-likelihood and the rest is the blob; the blob array takes ``blobs_dtype`` if given, else the dtype numpy
    infers from the first blob -- except text, which is stored as objects so that strings are never truncated --
    and length-1 trailing axes are dropped so that a single scalar blob per row yields a 1-d array."""
    tails = None
    try:
        if all(len(r) > 1 for r in rows) and len(rows):
            tails = [r[1:] for r in rows]
        elif any(len(r) > 1 for r in rows):
            tails = [r[1:] for r in rows if len(r) > 1]
    except TypeError:
        tails = None
    if not tails:
        return np.array([float(r) for r in rows]), None
    logl = np.array([float(r[0]) for r in rows])
    if owner is not None:
        owner.have_blobs = True
    dtype = blobs_dtype
    if dtype is None:
        try:
            dtype = np.atleast_1d(tails[0]).dtype
        except ValueError:          # ragged first blob
            dtype = np.dtype(object)
        if dtype.kind in ("U", "S"):
            dtype = np.dtype(object)
    blobs = np.array(tails, dtype=dtype)
    unit_axes = tuple(i for i in range(1, blobs.ndim) if blobs.shape[i] == 1)
    if unit_axes:
        blobs = np.squeeze(blobs, unit_axes)
    return logl, blobs
