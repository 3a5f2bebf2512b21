"""Vectorised MCMC kernels with the reference's dict-in / dict-out protocol (pocomc/mcmc.py):
``preconditioned_pcn`` (8-183), ``preconditioned_rwm`` (186-341), ``pcn`` (344-506), ``rwm``
(508-654).  One implementation drives all four: the particle state stays on the GPU for the whole
call, every step is a short chain of libpmc_b200 kernels
(noise -> proposal -> flow pull-back -> reparameterisation [+ device prior, same launch] -> [host prior / likelihood] ->
Metropolis update + scalar adaptation, one launch), each a programmatic dependent launch of its predecessor, and only
x' / logp' / logl' cross the PCIe bus because the user's log_likelihood is a host-side black box (x' through pinned memory, in
row chunks for large clouds: config.host_chunks).
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib, config, dist
from .flow import Flow

KIND_TPCN_FLOW, KIND_RWM_FLOW, KIND_TPCN, KIND_RWM = 0, 1, 2, 3
CTL_SIGMA, CTL_STEP, CTL_BEST, CTL_CNT, CTL_STOP, CTL_ACCEPT, CTL_CALLS, CTL_TRACK, CTL_NACC, CTL_MU = \
    0, 1, 2, 3, 4, 5, 6, 7, 8, 16

__all__ = ["preconditioned_pcn", "preconditioned_rwm", "pcn", "rwm", "McmcEngine"]


def _f64(a, dev):
    return torch.as_tensor(np.ascontiguousarray(a, dtype=np.float64)).to(dev)


class McmcEngine:
    """Device-resident state + buffers of one ``_mutate`` call."""

    def __init__(self, kind, state_dict, function_dict, option_dict):
        _lib.require_cuda()
        self.kind = kind
        self.use_flow = kind in (KIND_TPCN_FLOW, KIND_RWM_FLOW)
        self.tp = kind in (KIND_TPCN_FLOW, KIND_TPCN)
        dev = self.dev = torch.device("cuda", torch.cuda.current_device())
        # -- state copies (mcmc.py:31-41)
        x = np.copy(state_dict.get('x'))
        self.n, self.d = n, d = x.shape
        self.u = _f64(state_dict.get('u'), dev)
        self.x = _f64(x, dev)
        self.logdetj = _f64(state_dict.get('logdetj'), dev)
        self.logl = _f64(state_dict.get('logl'), dev)
        self.logp = _f64(state_dict.get('logp'), dev)
        self.beta = float(state_dict.get('beta'))
        self.blobs = state_dict.get('blobs')
        self.have_blobs = self.blobs is not None
        # -- functions (mcmc.py:44-48)
        self.log_like = function_dict.get('loglike')
        self.log_prior = function_dict.get('logprior')
        self.loglike_device = function_dict.get('loglike_device')
        self.logprior_device = function_dict.get('logprior_device')
        if self.logprior_device is None and config.device_prior:
            # Sampler passes prior.logpdf (sampler.py:203); a pocomc_b200.Prior of norm/uniform factors has a device form
            owner = getattr(self.log_prior, '__self__', None)
            spec = owner.device_spec() if hasattr(owner, 'device_spec') else None
            if spec is not None:
                from .synthetic import DevicePrior
                self.logprior_device = DevicePrior(*spec)
        # the device prior rides in the reparameterisation kernel (one launch for u' -> x', logdetj', finite, logp')
        self.prior_fused = bool(config.fuse_prior and hasattr(self.logprior_device, '_params'))
        self.scaler = function_dict.get('scaler')
        if not hasattr(self.scaler, '_params'):          # e.g. the reference's own Reparameterize (drop-in use from its Sampler)
            from .scaler import Reparameterize
            self.scaler = Reparameterize.adopt(self.scaler)
        geometry = function_dict.get('theta_geometry' if self.use_flow else 'u_geometry')
        self.with_bc = (self.scaler.periodic is not None) or (self.scaler.reflective is not None)
        # -- options (mcmc.py:51-54)
        self.n_max = int(option_dict.get('n_max'))
        self.n_steps = int(option_dict.get('n_steps'))
        self.progress_bar = option_dict.get('progress_bar')
        self.sweep_events = option_dict.get('sweep_events')     # optional list: (start, stop) CUDA events per inverse sweep
        sigma = option_dict.get('proposal_scale')
        if self.tp:
            sigma = np.minimum(sigma, 0.99)
        # -- flow push: theta, logdetj_flow = flow.forward(u) through the f32 shim (mcmc.py:60, tools.py:336-341)
        self.module = None
        self.theta = self.ldjf = None
        if self.use_flow:
            flow = function_dict.get('flow')
            if not isinstance(flow, Flow):
                raise TypeError("pocomc_b200 MCMC kernels need a pocomc_b200.Flow (sm_100a kernels); "
                                f"got {type(flow).__name__}")
            self.module = flow.flow.ensure_cuda()
            self.theta = torch.empty((n, d), dtype=torch.float32, device=dev)
            self.ldjf = torch.empty(n, dtype=torch.float32, device=dev)
            self.module.sweep_into(self.u.float(), self.theta, self.ldjf, inverse=False)
            self.ldjf.neg_()
        # -- geometry (mcmc.py:63-68 / 237-238)
        self.nu = 0.0
        self.inv_t = None
        mu = np.zeros(d)
        if self.tp:
            mu = np.asarray(geometry.t_mean, dtype=np.float64)
            cov = np.asarray(geometry.t_cov, dtype=np.float64)
            self.nu = float(geometry.t_nu)
            self.inv_t = _f64(np.linalg.inv(cov).T, dev)
            self.chol_t = _f64(np.linalg.cholesky(cov).T, dev)
        else:
            self.chol_t = _f64(np.linalg.cholesky(np.asarray(geometry.normal_cov, dtype=np.float64)).T, dev)
        # -- controller block
        logl0, logp0, ldj0 = (np.asarray(state_dict.get(k), dtype=np.float64) for k in ('logl', 'logp', 'logdetj'))
        best = np.mean(logl0 + logp0) if self.tp else np.mean(logl0 + logp0 + ldj0)   # mcmc.py:70 / 243
        shard = option_dict.get('shard')
        if shard is not None and dist.is_active():
            # the plateau tracker starts from the mean over ALL particles (mcmc.py:70 / 243): the caller that holds the
            # full arrays hands it over (bit-identical to a single-GPU run); otherwise rank-ordered sum of shard sums
            if option_dict.get('best0') is not None:
                best = float(option_dict['best0'])
            else:
                tot = logl0 + logp0 if self.tp else logl0 + logp0 + ldj0
                best = float(dist.allreduce_sum_det(torch.tensor([np.sum(tot)], dtype=torch.float64, device=dev))[0].item()) / shard[1]
        ctl = np.zeros(CTL_MU + d)
        ctl[CTL_SIGMA], ctl[CTL_BEST] = sigma, best
        ctl[CTL_MU:] = mu
        self.ctl = _f64(ctl, dev)
        self.ctl_host = _lib.pinned('ctl', CTL_MU + d, torch.float64)
        # -- per-step buffers
        f64 = dict(dtype=torch.float64, device=dev)
        self.g = torch.empty(n, **f64) if self.tp else None
        self.z = torch.empty((n, d), **f64)
        self.r = torch.empty(n, **f64)
        self.prop64 = torch.empty((n, d), **f64)
        self.prop32 = torch.empty((n, d), dtype=torch.float32, device=dev) if self.use_flow else None
        self.m_cur = torch.empty(n, **f64) if self.tp else None
        self.m_prop = torch.empty(n, **f64) if self.tp else None
        self.u_p32 = torch.empty((n, d), dtype=torch.float32, device=dev) if self.use_flow else None
        self.ldjf_p = torch.empty(n, dtype=torch.float32, device=dev) if self.use_flow else None
        self.u_p = torch.empty((n, d), **f64)
        self.x_p = torch.empty((n, d), **f64)
        self.ldj_p = torch.empty(n, **f64)
        self.finite = torch.empty(n, dtype=torch.uint8, device=dev)
        self.logl_p = torch.empty(n, **f64)
        self.logp_p = torch.empty(n, **f64)
        self.alpha = torch.empty(n, **f64)
        self.partials = torch.empty(int(_lib.load().pmc_mh_partials_size(n, d)), **f64)
        self.ticket = torch.zeros(1, dtype=torch.int32, device=dev)      # last-block-done counter of the fused accept + adapt launch
        # pinned staging
        self.h_x = _lib.pinned('x', (n, d), torch.float64)
        self.h_fin = _lib.pinned('fin', n, torch.uint8)
        self.h_ll = _lib.pinned('ll', (2, n), torch.float64)
        self.h_noise = _lib.pinned('noise', n * (d + 2), torch.float64) if config.rng_mode == "host" else None
        self.rng_mode = config.rng_mode
        self.mean_mode = config.resolved_mean_mode()
        # particle sharding: option_dict['shard'] = (global offset of row 0, global particle count, blocks per rank)
        self.sharded = shard is not None and dist.is_active()
        self.row_offset, self.n_global, self.shard_blocks = (shard if shard is not None else (0, n, None))
        self._peer = None
        if self.sharded:
            self.mean_mode = 0
            self._peer = dist.peer_exchange(self.shard_blocks, d + 4)      # None: all-gather path (gloo / shared GPU)
        self.seed = 0
        if self.rng_mode == "device":     # one draw from the host stream keys the Philox counters (same on every rank)
            self.seed = int(option_dict['seed']) if option_dict.get('seed') is not None else int(np.random.randint(0, 2 ** 62))
        self.n_calls = 0
        self.launches = 0       # libpmc_b200 kernels launched by this engine (bench.py's gpu_launches)
        self.step = 0
        self.sigma = float(sigma)
        self.accept = 0.0
        self.stop = False
        self._propose = self._scaler_inverse = self._finalize = None       # pre-bound libpmc_b200 calls (built on first use)
        self._rng_fill = self._sweep = self._logprior = None
        self._download = self._read_ctl = None
        self._stream = torch.cuda.current_stream()     # the engine lives for one kernel call on the caller's stream; looking it up costs ~20 us per step
        self._accept = {}
        self._accept_fused = {}
        self.sc, self._sc_keep, _ = self.scaler._params(True)
        if self.with_bc and self._sc_keep["bc"] is not None:
            self.sc.bc = _lib.ptr(self._sc_keep["bc"])

    # -- one step ---------------------------------------------------------------------------------
    def draw_noise(self):
        """N gammas then N*D normals (mcmc.py:80,85 / 253), host stream or Philox."""
        n, d = self.n, self.d
        if self.rng_mode == "host":
            # sharded: every rank walks the same global stream (identical seeds) and keeps its own rows, so the
            # trajectory of a particle does not depend on the number of GPUs
            hn = self.h_noise.numpy()
            ng, lo = (self.n_global, self.row_offset) if self.sharded else (n, 0)
            if self.tp:
                hn[:n] = np.random.standard_gamma((d + self.nu) / 2, size=ng)[lo:lo + n]   # == gamma(a, s_k)/s_k draw for draw
                self.g.copy_(self.h_noise[:n], non_blocking=True)
            hn[n:n + n * d] = np.random.randn(ng, d)[lo:lo + n].reshape(-1)
            self.z.copy_(self.h_noise[n:n + n * d].view(n, d), non_blocking=True)
        else:
            if self._rng_fill is None:       # the step counter is read on the device (ctl[STEP] + 1): nothing changes per call
                self._rng_fill = _lib.bind("pmc_rng_fill_ctl", C.c_uint64(self.seed), _lib.ptr(self.ctl), int(self.row_offset),
                                           (d + self.nu) / 2 if self.tp else 0.0, _lib.ptr(self.g), _lib.ptr(self.z),
                                           _lib.ptr(self.r), n, d)
            self._rng_fill()

    def propose(self):
        if self._propose is None:          # every buffer of the engine is fixed: bind the argument list once
            n, d = self.n, self.d
            pos = self.theta if self.use_flow else self.u
            if self.tp:
                self._propose = _lib.bind("pmc_tpcn_propose", 1 if self.use_flow else 0, _lib.ptr(pos), _lib.ptr(self.ctl),
                                          _lib.ptr(self.inv_t), _lib.ptr(self.chol_t), self.nu, _lib.ptr(self.g),
                                          _lib.ptr(self.z), _lib.ptr(self.prop64), _lib.ptr(self.prop32),
                                          _lib.ptr(self.m_cur), _lib.ptr(self.m_prop), n, d)
            else:
                self._propose = _lib.bind("pmc_rwm_propose", 1 if self.use_flow else 0, _lib.ptr(pos), _lib.ptr(self.ctl),
                                          _lib.ptr(self.chol_t), _lib.ptr(self.z), _lib.ptr(self.prop64),
                                          _lib.ptr(self.prop32), n, d)
        self._propose()

    def pull_back(self):
        """theta' -> u' (flow.inverse, mcmc.py:88) -> x', logdetj' (+ boundary conditions, :91-97)."""
        n, d = self.n, self.d
        if self.use_flow:
            if self.sweep_events is not None:
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
            if self._sweep is None:
                self._sweep = self.module.bind_sweep(self.prop32, self.u_p32, self.ldjf_p, inverse=True)
            self._sweep()
            if self.sweep_events is not None:
                e1.record()
                self.sweep_events.append((e0, e1))
            src, is32 = self.u_p32, 1
        else:
            src, is32 = self.prop64, 0
        if self._scaler_inverse is None:
            if self.prior_fused:
                self._prior_keep = kind, loc, scale = self.logprior_device._params(self.dev)
                self._scaler_inverse = _lib.bind("pmc_scaler_inverse_prior", is32, _lib.ptr(src), C.byref(self.sc), _lib.ptr(kind),
                                                 _lib.ptr(loc), _lib.ptr(scale), _lib.ptr(self.u_p), _lib.ptr(self.x_p),
                                                 _lib.ptr(self.ldj_p), _lib.ptr(self.finite), _lib.ptr(self.logp_p), n, d)
            else:
                self._scaler_inverse = _lib.bind("pmc_scaler_inverse", is32, _lib.ptr(src), C.byref(self.sc), _lib.ptr(self.u_p),
                                                 _lib.ptr(self.x_p), _lib.ptr(self.ldj_p), _lib.ptr(self.finite), n, d)
        self._scaler_inverse()

    def _bind_host_io(self):
        """Pre-bound staging calls of the host-likelihood step: chunked download of x' (+ finite flags) with one event per
        chunk, upload of logl' / logp', download of the controller block."""
        n, d = self.n, self.d
        k = config.host_chunks
        if k <= 0:
            k = 4 if n * d * 8 >= (16 << 20) else 1       # measured: 10 000 x 32-D loses with any split, 50 000 x 50-D gains 4 % with 4
        if self.have_blobs:
            k = 1
        k = max(1, min(int(k), n))
        per = (n + k - 1) // k
        self._chunks = [(a, min(a + per, n)) for a in range(0, n, per)]
        k = len(self._chunks)
        self._events = _lib.Events.get(k)
        self._download = _lib.bind("pmc_download_rows", _lib.ptr(self.x_p), _lib.ptr(self.h_x), _lib.ptr(self.finite),
                                   _lib.ptr(self.h_fin), n, d, k, self._events.handles, kernel=False)
        self._up_logl = _lib.bind("pmc_memcpy_async", _lib.ptr(self.logl_p), _lib.ptr(self.h_ll[0]), n * 8, kernel=False)
        self._up_logp = _lib.bind("pmc_memcpy_async", _lib.ptr(self.logp_p), _lib.ptr(self.h_ll[1]), n * 8, kernel=False)

    def evaluate_host(self):
        """Host black boxes on the finite rows only (mcmc.py:100-121).  When the prior is a
        ``pocomc_b200.Prior`` of frozen scipy norm / uniform factors it is evaluated on the GPU
        (config.device_prior) and only the likelihood crosses the PCIe bus.  x' comes down in row chunks
        (config.host_chunks); the callables see each chunk as soon as it has landed."""
        n = self.n
        if self.logprior_device is not None and not self.prior_fused:
            if self._logprior is None:
                bind = getattr(self.logprior_device, "bind", None)
                self._logprior = bind(self.x_p, self.finite, self.logp_p) if bind is not None else \
                    (lambda: self.logprior_device(self.x_p, self.finite, self.logp_p))
            self._logprior()                                                # also clears finite where logp' is not finite
        if self._download is None:
            self._bind_host_io()
        self._download()
        x_all = self.h_x.numpy()
        mask_all = self.h_fin.numpy().view(np.bool_)
        hl = self.h_ll.numpy()
        host_prior = self.logprior_device is None
        calls = 0
        blobs_p = None
        for (a, b), wait in zip(self._chunks, self._events.wait):
            wait()
            x_p, mask = x_all[a:b], mask_all[a:b]
            all_rows = bool(mask.all())
            if host_prior:
                logp_p = hl[1, a:b]
                if all_rows:
                    logp_p[:] = self.log_prior(x_p)
                else:
                    logp_p.fill(-np.inf)
                    logp_p[mask] = self.log_prior(x_p[mask])
                ok = np.isfinite(logp_p)
                if not ok.all():
                    mask = mask & ok
                    all_rows = False
            logl_p = hl[0, a:b]
            if all_rows and not self.have_blobs:
                logl_p[:], _ = self.log_like(x_p)
            else:
                logl_p.fill(-np.inf)
                if self.have_blobs:                                          # one chunk: a = 0, b = n
                    blobs_p = np.empty(n, dtype=np.dtype((self.blobs[0].dtype, self.blobs[0].shape)))
                    logl_p[mask], blobs_p[mask] = self.log_like(x_p[mask])
                elif mask.any() or len(self._chunks) == 1:               # a single chunk keeps the reference's call on zero rows
                    logl_p[mask], _ = self.log_like(x_p[mask])
            calls += (b - a) if all_rows else int(np.sum(mask))
        if host_prior:
            self._up_logp()
        self._up_logl()
        return calls, blobs_p

    def evaluate_device(self):
        """Opt-in device prior / likelihood (synthetic benchmarks): no PCIe traffic."""
        if not self.prior_fused:
            self.logprior_device(self.x_p, self.finite, self.logp_p)
        self.loglike_device(self.x_p, self.finite, self.logl_p)
        return None, None

    def accept_and_adapt(self, calls):
        n, d = self.n, self.d
        if self.rng_mode == "host":
            hn = self.h_noise.numpy()
            ng, lo = (self.n_global, self.row_offset) if self.sharded else (n, 0)
            hn[n + n * d:] = np.random.rand(ng)[lo:lo + n]                       # mcmc.py:137
            self.r.copy_(self.h_noise[n + n * d:], non_blocking=True)
        key = calls is None
        if not self.sharded:
            # one launch: Metropolis update + block partials, and the block that finishes last adapts sigma / mu and
            # applies the plateau rule (no-op once the stop flag is set)
            if key not in self._accept_fused:
                self._accept_fused[key] = _lib.bind(
                    "pmc_mh_accept_finalize", self.kind, self.beta, self.nu, _lib.ptr(self.theta), _lib.ptr(self.u),
                    _lib.ptr(self.x), _lib.ptr(self.logdetj), _lib.ptr(self.logl), _lib.ptr(self.logp), _lib.ptr(self.ldjf),
                    _lib.ptr(self.prop64), _lib.ptr(self.u_p), _lib.ptr(self.x_p), _lib.ptr(self.ldj_p),
                    _lib.ptr(self.logl_p), _lib.ptr(self.logp_p), _lib.ptr(self.ldjf_p), _lib.ptr(self.m_cur),
                    _lib.ptr(self.m_prop), _lib.ptr(self.r), _lib.ptr(self.finite) if calls is None else None,
                    _lib.ptr(self.alpha), _lib.ptr(self.partials), _lib.ptr(self.ctl), _lib.ptr(self.ticket), self.mean_mode,
                    self.n_steps, self.n_max, n, d)
            self._accept_fused[key]()
            return
        if self._peer is not None:
            # sharded, one GPU per rank: accept + exchange of the block partials through peer memory + adaptation, one launch
            if key not in self._accept_fused:
                self._accept_fused[key] = _lib.bind(
                    "pmc_mh_accept_finalize_p2p", self.kind, self.beta, self.nu, _lib.ptr(self.theta), _lib.ptr(self.u),
                    _lib.ptr(self.x), _lib.ptr(self.logdetj), _lib.ptr(self.logl), _lib.ptr(self.logp), _lib.ptr(self.ldjf),
                    _lib.ptr(self.prop64), _lib.ptr(self.u_p), _lib.ptr(self.x_p), _lib.ptr(self.ldj_p),
                    _lib.ptr(self.logl_p), _lib.ptr(self.logp_p), _lib.ptr(self.ldjf_p), _lib.ptr(self.m_cur),
                    _lib.ptr(self.m_prop), _lib.ptr(self.r), _lib.ptr(self.finite) if calls is None else None,
                    _lib.ptr(self.alpha), _lib.ptr(self.partials), _lib.ptr(self.ctl), _lib.ptr(self.ticket),
                    self.n_steps, self.n_max, n, d, self._peer, self.n_global)
            self._accept_fused[key]()
            return
        key = calls is None
        if key not in self._accept:
            self._accept[key] = _lib.bind(
                "pmc_mh_accept_update", self.kind, self.beta, self.nu, _lib.ptr(self.theta), _lib.ptr(self.u),
                _lib.ptr(self.x), _lib.ptr(self.logdetj), _lib.ptr(self.logl), _lib.ptr(self.logp), _lib.ptr(self.ldjf),
                _lib.ptr(self.prop64), _lib.ptr(self.u_p), _lib.ptr(self.x_p), _lib.ptr(self.ldj_p),
                _lib.ptr(self.logl_p), _lib.ptr(self.logp_p), _lib.ptr(self.ldjf_p), _lib.ptr(self.m_cur),
                _lib.ptr(self.m_prop), _lib.ptr(self.r), _lib.ptr(self.finite) if calls is None else None,
                _lib.ptr(self.alpha), _lib.ptr(self.partials), n, d)
        self._accept[key]()
        if self.sharded:      # rank-ordered all-gather of the block partials, summed in fixed order by every rank
            parts = dist.gather_blocks(self.partials.view(-1, d + 4), self.shard_blocks)
            _lib.call("pmc_mcmc_finalize", self.kind, _lib.ptr(self.ctl), _lib.ptr(parts), parts.shape[0], _lib.ptr(self.theta),
                      self.mean_mode, self.n_steps, self.n_max, self.n_global, d)
            return

    def read_controller(self):
        if self._read_ctl is None:
            self._read_ctl = (_lib.bind("pmc_memcpy_async", _lib.ptr(self.ctl_host), _lib.ptr(self.ctl), self.ctl.numel() * 8, kernel=False),
                              _lib.bind("pmc_stream_synchronize", kernel=False))
        self._read_ctl[0]()
        self._read_ctl[1]()
        c = self.ctl_host.numpy()
        self.sigma, self.accept = float(c[CTL_SIGMA]), float(c[CTL_ACCEPT])
        self.step, self.stop = int(c[CTL_STEP]), bool(c[CTL_STOP] != 0.0)
        if self.stop and self._peer is not None and dist.peer_exchange_error():
            raise RuntimeError("pocomc_b200: a peer rank did not publish its block partials within 20 s (peer-memory exchange)")
        return c

    def reset_controller(self):
        """Start another ``_mutate``-sized run from the current state: step, plateau counter and
        stop flag cleared, sigma and mu kept (what successive Sampler._mutate calls at one beta do)."""
        self.ctl[CTL_STEP:CTL_STOP + 1] = torch.tensor([0.0, float((self.logl + self.logp).mean().item()) if self.tp else
                                                        float((self.logl + self.logp + self.logdetj).mean().item()), 0.0, 0.0],
                                                       dtype=torch.float64, device=self.dev)
        self.step, self.stop = 0, False

    def run(self):
        self.loop()
        return self.results()

    def _steps_before_check(self, c):
        """How many steps may be queued before the controller is read back.  Exactness does not depend on it (queued
        steps become no-ops once the stop flag is set); this only avoids queueing work that the plateau rule of
        mcmc.py:170-180 would discard: the counter grows by at most one per step."""
        if c is None:
            return 1
        sigma = max(abs(float(c[CTL_SIGMA])), 1e-300)
        ratio = (2.38 / np.sqrt(self.d)) / sigma
        if self.kind == KIND_RWM_FLOW:
            ratio = min(1.0, ratio)
        left = min(self.n_steps * ratio ** 2 - float(c[CTL_CNT]), self.n_max - float(c[CTL_STEP]))
        return int(max(1, min(8, np.floor(left))))

    def _after_step(self, c, step_calls, blobs_p):
        if self.have_blobs:
            acc = (self.r < self.alpha).cpu().numpy()
            self.blobs[acc] = blobs_p[acc]                                   # mcmc.py:148-149
        if self.progress_bar is not None:                                    # mcmc.py:159-167
            self.progress_bar.update_stats(dict(
                calls=self.progress_bar.info['calls'] + step_calls, acc=self.accept, steps=self.step,
                logP=float(c[CTL_TRACK]) if self.tp else float((self.logl + self.logp).mean().item()),
                eff=self.sigma / (2.38 / np.sqrt(self.d))))

    def loop(self):
        """MCMC steps until the plateau rule or n_max fires (mcmc.py:72-180); state stays on the GPU."""
        device_eval = self.loglike_device is not None and self.logprior_device is not None and not self.have_blobs
        fused = not self.sharded or self._peer is not None       # accept + (exchange +) adaptation in one stop-guarded launch
        per_step = (1 if self.rng_mode == "device" else 0) + 1 + (1 if self.use_flow else 0) + 1 + (1 if fused else 2)
        if device_eval and self.rng_mode == "device" and fused:
            # nothing of a step needs the host: queue several steps per controller read-back.  Every kernel of a step
            # reads its scalars (sigma, mu, step) from the device controller block; the fused accept + adapt launch is
            # a no-op after the stop flag is set, so the state stops changing exactly where the reference stops.
            c = None
            while True:
                k = self._steps_before_check(c)
                for _ in range(k):
                    self.draw_noise()
                    self.propose()
                    self.pull_back()
                    self.evaluate_device()
                    self.accept_and_adapt(None)
                done_before = self.step
                c = self.read_controller()
                self.launches += (per_step + (1 if self.prior_fused else 2)) * k
                step_calls = int(c[CTL_CALLS]) - self.n_calls
                self.n_calls = int(c[CTL_CALLS])
                if self.step > done_before or self.stop:
                    self._after_step(c, step_calls, None)
                if self.stop:
                    break
            return
        while True:
            self.draw_noise()
            self.propose()
            self.pull_back()
            self.launches += per_step + ((1 if self.prior_fused else 2) if device_eval else 0) + \
                (1 if (not device_eval and self.logprior_device is not None and not self.prior_fused) else 0)
            if device_eval:
                calls, blobs_p = self.evaluate_device()
            else:
                calls, blobs_p = self.evaluate_host()
                self.n_calls += calls
            self.accept_and_adapt(calls)
            c = self.read_controller()
            if device_eval:
                step_calls = int(c[CTL_CALLS]) - self.n_calls
                self.n_calls = int(c[CTL_CALLS])
            else:
                step_calls = calls
            self._after_step(c, step_calls, blobs_p)
            if self.stop:
                break

    def results(self):
        """Download the final state in the reference's result-dict layout (mcmc.py:182-183)."""
        out = dict(u=self.u.cpu().numpy(), x=self.x.cpu().numpy(), logdetj=self.logdetj.cpu().numpy(),
                   logl=self.logl.cpu().numpy(), logp=self.logp.cpu().numpy(), blobs=self.blobs,
                   efficiency=self.sigma, accept=self.accept, steps=self.step, calls=self.n_calls,
                   proposal_scale=self.sigma)
        if self.sharded and self.loglike_device is not None and self.logprior_device is not None and not self.have_blobs:
            out["calls_global"] = True           # counted from the all-gathered partials: already the sum over ranks
        return out


@torch.no_grad()
def preconditioned_pcn(state_dict: dict, function_dict: dict, option_dict: dict):
    """Doubly preconditioned Crank-Nicolson in flow-latent space (mcmc.py:8-183)."""
    return McmcEngine(KIND_TPCN_FLOW, state_dict, function_dict, option_dict).run()


@torch.no_grad()
def preconditioned_rwm(state_dict: dict, function_dict: dict, option_dict: dict):
    """Preconditioned random-walk Metropolis in flow-latent space (mcmc.py:186-341)."""
    return McmcEngine(KIND_RWM_FLOW, state_dict, function_dict, option_dict).run()


def pcn(state_dict: dict, function_dict: dict, option_dict: dict):
    """t-preconditioned Crank-Nicolson in u space (mcmc.py:344-506)."""
    return McmcEngine(KIND_TPCN, state_dict, function_dict, option_dict).run()


def rwm(state_dict: dict, function_dict: dict, option_dict: dict):
    """Random-walk Metropolis in u space (mcmc.py:508-654)."""
    return McmcEngine(KIND_RWM, state_dict, function_dict, option_dict).run()
