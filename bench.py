#!/usr/bin/env python
"""bench.py -- particle-steps/s of pocoMC's flow-preconditioned MCMC hot path (BASELINE.json).

One bench "step" = one `_mutate`-sized call of the t-preconditioned Crank-Nicolson kernel (pocomc/mcmc.py:8-183):
MCMC_STEPS Metropolis steps over every particle with the plateau rule disabled; particle-steps/s = particles x MCMC
steps / time.  `--config` selects the workload (default 1 = the configuration BASELINE's metric is quoted on):

  0  10-D Rosenbrock,            U(-10,10)^10,  1 000 particles            (README example; the reference's CPU case)
  1  32-D correlated Gaussian,   N(0,3^2)^32,  10 000 particles            (headline; flow-preconditioned MCMC only)
  2  50-D bimodal mixture,       U(-10,10)^50, 50 000 particles
  3  100-D Rosenbrock,           U(-10,10)^100, 200 000 particles over 4 GPUs = 50 000 per GPU
  4  200-D Neal funnel,          U(-30,30)^200, 1 000 000 particles over 8 GPUs = 125 000 per GPU
  5  32-D Rosenbrock,            U(-10,10)^32, 10 000 particles            (north_star's literal target line: probit scaler path)

all with flow 'maf6' trained for 200 optimiser steps on the initial cloud, Student-t geometry fitted on the latent
cloud, beta = 1.  At N GPUs every rank holds the per-GPU particle count (weak scaling); `--scaling strong` keeps the
TOTAL fixed (the config's own multi-GPU total, else 64 x its particle count, rounded to whole tile waves: 606 208 particles
at config 1) and gives every rank total / N.

  value : device-resident arm -- state in HBM, Philox noise and the synthetic prior/likelihood evaluated on the GPU, no
          host traffic inside the timed region.
  e2e   : the reference-facing call `pocomc_b200.mcmc.preconditioned_pcn(state_dict, function_dict, option_dict)` with
          HOST numpy buffers, the likelihood as a host black box (x' D2H and logl' H2D every MCMC step).
  aux   : (config 1, N = 1) untimed-region extras: tcgen05 dense forward, a measured TF32 GEMM peak, one Flow.fit
          optimiser step, and a FULL Sampler.run() of configs[0] against the unmodified reference's logZ.
  --impl reference / cpu_baseline : the UNMODIFIED reference (`oracle/_ref/pocomc`, staged by oracle/make_ref.sh) --
          its own `pocomc.mcmc.preconditioned_pcn`, `Flow`, `Reparameterize`, `Geometry` -- over `oracle/zuko` (the
          restatement of its one missing third-party dependency), on all host threads, on a bounded sample of the same
          workload.  Falls back to the in-repo oracle port (kind "port") only if oracle/_ref is absent.
"""
from __future__ import annotations

import argparse
import json
import math
import os
import re
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FLOW = "maf6"
TRAIN_EPOCHS = 20        # x 10 batches of 512 (train split = 5000 rows) = 200 optimiser steps
L2_FLUSH_BYTES = 256 << 20

# name, D, particles per GPU, GPUs the config is defined on, MCMC steps per bench step, CPU-sample particles
CONFIGS = {
    0: dict(name="10-D Rosenbrock, n_particles=1000", d=10, n=1000, gpus=1, steps=50, like="rosen", prior=("uniform", -10.0, 20.0), cpu_n=1000),
    1: dict(name="32-D correlated Gaussian, n_particles=10000", d=32, n=10_000, gpus=1, steps=50, like="gauss", prior=("norm", 0.0, 3.0), cpu_n=10_000),
    2: dict(name="50-D bimodal Gaussian mixture, n_particles=50000", d=50, n=50_000, gpus=1, steps=10, like="mixture", prior=("uniform", -10.0, 20.0), cpu_n=2000),
    3: dict(name="100-D Rosenbrock, n_particles=200000 over 4 GPUs", d=100, n=50_000, gpus=4, steps=5, like="rosen", prior=("uniform", -10.0, 20.0), cpu_n=500),
    4: dict(name="200-D funnel, n_particles=1000000 over 8 GPUs", d=200, n=125_000, gpus=8, steps=3, like="funnel", prior=("uniform", -30.0, 60.0), cpu_n=200),
    5: dict(name="32-D Rosenbrock, U(-10,10), n_particles=10000 (north_star target line)", d=32, n=10_000, gpus=1, steps=50, like="rosen", prior=("uniform", -10.0, 20.0), cpu_n=10_000),
}


# ---------------------------------------------------------------------------------------------
# workload (numpy / scipy only: shared by every arm)
# ---------------------------------------------------------------------------------------------
class Workload:
    def __init__(self, cfg, n, seed=0, row_offset=0):
        from scipy.stats import norm, uniform
        from pocomc_b200 import synthetic as S
        self.cfg, self.n, self.d = cfg, n, cfg["d"]
        d = self.d
        kind, a, b = cfg["prior"]
        self.dists = [(norm if kind == "norm" else uniform)(a, b)] * d
        self.prior_kind = 0 if kind == "norm" else 1
        self.prior_loc, self.prior_scale = a, b
        self.bounds = np.array([dd.support() for dd in self.dists])
        self.like = dict(gauss=lambda: S.CorrelatedGaussian(d), rosen=lambda: S.Rosenbrock(), mixture=lambda: S.GaussianMixture(),
                         funnel=lambda: S.Funnel())[cfg["like"]]()
        rng = np.random.default_rng(seed)
        self.prior_samples = np.stack([dd.rvs(size=4096, random_state=rng) for dd in self.dists], axis=1)   # scaler.fit input (same on all ranks)
        self.x0 = self.cloud(n, np.random.default_rng([seed, 1, row_offset]))

    def cloud(self, n, rng):
        """a crude posterior-like cloud to start the chains from (throughput does not depend on it being exact)"""
        d, like = self.d, self.cfg["like"]
        if like == "gauss":
            return rng.normal(size=(n, d)) @ np.linalg.cholesky(self.like.cov).T
        if like == "mixture":
            sign = np.where(rng.random(n) < 0.5, 1.0, -1.0)[:, None]
            return sign * self.like.p0 + self.like.p1 * rng.normal(size=(n, d))
        if like == "rosen":
            x = np.empty((n, d))
            x[:, ::2] = rng.normal(0.8, 0.3, size=(n, (d + 1) // 2))
            x[:, 1::2] = x[:, ::2][:, :d // 2] ** 2 + rng.normal(0.0, 0.15, size=(n, d // 2))
            return np.clip(x, -9.9, 9.9)
        x0 = rng.normal(0.0, 1.5, size=n)
        x = rng.normal(size=(n, d)) * np.exp(0.5 * x0)[:, None]
        x[:, 0] = x0
        return np.clip(x, -29.0, 29.0)

    def loglike(self, x):
        return self.like(x)

    def logprior(self, x):
        out = np.zeros(len(x))
        for i, dd in enumerate(self.dists):
            out += dd.logpdf(x[:, i])
        return out


def clocks_sampler(stop_evt, out):
    q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    try:
        p = subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "100",
                              "-i", os.environ.get("LOCAL_RANK", "0")], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
    except OSError:
        return

    def reader():
        for line in p.stdout:
            out.append(line.strip())
    t = threading.Thread(target=reader, daemon=True)
    t.start()
    stop_evt.wait()
    p.terminate()
    t.join(timeout=2)


def summarise_clocks(lines):
    sm, mx, reasons = [], [], set()
    names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
    for ln in lines:
        f = [s.strip() for s in ln.split(",")]
        if len(f) < 9:
            continue
        try:
            sm.append(float(f[1])); mx.append(float(f[2]))
        except ValueError:
            continue
        for name, v in zip(names, f[5:9]):
            if v.lower().startswith("active"):
                reasons.add(name)
    if not sm:
        return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
    return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(np.max(mx)), "reasons": sorted(reasons), "samples": len(sm)}


def workload_config(cfg_id, cfg, n_local, n_gpus, mcmc_steps, scaling):
    return {"workload": f"configs[{min(cfg_id, 4)}{'*' if cfg_id == 5 else ''}]: {cfg['name']}; flow-preconditioned MCMC only (tpCN, {FLOW}, beta=1)",
            "config_id": cfg_id, "n_particles_per_gpu": n_local, "n_particles_total": n_local * n_gpus, "n_dim": cfg["d"], "flow": FLOW,
            "mcmc_steps_per_bench_step": mcmc_steps, "kernel": "preconditioned_pcn",
            "parallelism": (f"particle-shard x{n_gpus} ({scaling})" if n_gpus > 1 else "single GPU") +
                           (f"; the config is defined on {cfg['gpus']} GPUs: this run holds one GPU's shard" if cfg["gpus"] > n_gpus else ""),
            "l2": "the MCMC state of one shard is L2-resident within a bench step by design; L2 flushed (256 MiB write) between bench steps"}


# ---------------------------------------------------------------------------------------------
# reference arm / cpu baseline
# ---------------------------------------------------------------------------------------------
def reference_problem(wl, flow_params=None, threads=None, n_sub=None):
    """The same `_mutate` call on the UNMODIFIED reference (oracle/_ref) over oracle/zuko; returns (run(steps) -> seconds, kind)."""
    import torch
    ref = os.path.join(ROOT, "oracle", "_ref")
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    if threads:
        torch.set_num_threads(threads)
    n = wl.n if n_sub is None else min(n_sub, wl.n)
    x0 = wl.x0[:n]
    if not os.path.isdir(os.path.join(ref, "pocomc")):
        return _port_problem(wl, x0, flow_params), "port"
    sys.path.insert(0, ref)
    import pocomc as rpc                                     # the reference, not this repo
    from pocomc import mcmc as rmcmc
    from pocomc.geometry import Geometry as RGeometry
    from pocomc.scaler import Reparameterize as RReparameterize
    from pocomc.tools import flow_numpy_wrapper as rwrap
    assert os.path.abspath(rpc.__file__).startswith(ref), rpc.__file__
    scaler = RReparameterize(wl.d, bounds=wl.bounds)
    scaler.fit(wl.prior_samples)
    u0 = scaler.forward(x0)
    ldj0 = scaler.inverse(u0)[1]
    torch.manual_seed(0)
    flow = rpc.Flow(wl.d, FLOW)
    if flow_params is not None:
        with torch.no_grad():
            for p_, v in zip(flow.flow.parameters(), flow_params):
                p_.copy_(torch.as_tensor(v).reshape(p_.shape))
    else:
        flow.fit(torch.tensor(u0, dtype=torch.float32), validation_split=0.5, epochs=TRAIN_EPOCHS, batch_size=512, patience=10 ** 6,
                 annealing=False)
    theta0 = rwrap(flow).forward(u0)[0]
    geo = RGeometry()
    geo.fit(theta0.astype(np.float64))
    state = dict(u=u0, x=x0, logdetj=ldj0, logl=wl.loglike(x0), logp=wl.logprior(x0), beta=1.0, blobs=None)
    fd = dict(loglike=lambda x: (wl.loglike(x), None), logprior=wl.logprior, scaler=scaler, flow=flow, theta_geometry=geo, u_geometry=geo)

    def run(mcmc_steps):
        od = dict(n_max=mcmc_steps, n_steps=10 ** 9, progress_bar=None, proposal_scale=2.38 / wl.d ** 0.5)
        t0 = time.perf_counter()
        res = rmcmc.preconditioned_pcn({k: (v.copy() if isinstance(v, np.ndarray) else v) for k, v in state.items()}, fd, od)
        dt = time.perf_counter() - t0
        assert res["steps"] == mcmc_steps
        return dt
    run.n = n
    return run, "reference"


def _port_problem(wl, x0, flow_params):
    import torch
    import flow_ref as F
    import smc_ref as O
    scaler = O.scaler_fit(wl.prior_samples, wl.bounds[:, 0], wl.bounds[:, 1])
    u0 = O.scaler_forward(x0, scaler)
    _, ldj0 = O.scaler_inverse(u0, scaler)
    torch.manual_seed(0)
    flow = F.make_flow(wl.d, FLOW)
    if flow_params is not None:
        F.load_params(flow, flow_params)
    else:
        F.fit(flow, torch.tensor(u0, dtype=torch.float32), validation_split=0.5, epochs=TRAIN_EPOCHS, batch_size=512, patience=10 ** 6)
    nf = F.NumpyFlow(flow)
    theta0, _ = nf.forward(u0)
    t_mean, t_cov, t_nu = O.fit_mvstud(theta0.astype(np.float64))
    geo = dict(t_mean=t_mean, t_cov=t_cov, t_nu=t_nu if np.isfinite(t_nu) else 1e6)
    state = dict(u=u0, x=x0, logdetj=ldj0, logl=wl.loglike(x0), logp=wl.logprior(x0), beta=1.0)

    def run(mcmc_steps):
        t0 = time.perf_counter()
        res = O.mcmc_kernel("tpcn_flow", state, wl.loglike, wl.logprior, scaler, geo,
                            dict(n_max=mcmc_steps, n_steps=10 ** 9, proposal_scale=2.38 / wl.d ** 0.5), flow=nf)
        assert res["steps"] == mcmc_steps
        return time.perf_counter() - t0
    run.n = len(x0)
    return run


def run_reference(args):
    if int(os.environ.get("RANK", "0")) != 0:
        return
    cfg = CONFIGS[args.config]
    threads = os.cpu_count() or 1
    wl = Workload(cfg, cfg["cpu_n"])
    np.random.seed(0)
    run, kind = reference_problem(wl, threads=threads)
    for _ in range(args.warmup):
        run(1)
    times = [run(1) for _ in range(args.steps)]
    total = float(np.sum(times))
    value = run.n * args.steps / total
    sample = (f"1 tpCN step x {run.n} particles per bench step"
              + (f" (a {run.n}-particle sample of the {cfg['n']}-particle shard; per-particle cost is size-independent)" if run.n < cfg["n"] else "")
              + ("; unmodified pocomc.mcmc.preconditioned_pcn + pocomc.Flow over oracle/zuko" if kind == "reference" else "; oracle port"))
    line = dict(impl="reference", metric="particle-steps/sec", value=value, unit="particle-steps/s", n_gpus=args.gpus,
                steps=args.steps, warmup=args.warmup, ms_per_step=1e3 * total / args.steps, higher_is_better=True,
                scaling=args.scaling, vs_baseline=None, dtype="f32 flow / f64 SMC state", data="synthetic",
                config=workload_config(args.config, cfg, cfg["n"], 1, 1, args.scaling),
                cpu_baseline=dict(value=value, unit="particle-steps/s", cores=threads, kind=kind, sample=sample),
                e2e=dict(value=value, unit="particle-steps/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0))
    print(json.dumps(line))


# ---------------------------------------------------------------------------------------------
# roofline helpers
# ---------------------------------------------------------------------------------------------
def committed_traffic(kernel_name, n, n_dim):
    """dram bytes per launch of the dominant kernel from the committed ncu --set full captures (profiles/roofline_traffic.json:
    kernel-name regex -> bytes, source file).  None (and a note) when no committed capture matches the kernel that ran."""
    try:
        table = json.load(open(os.path.join(ROOT, "profiles", "roofline_traffic.json")))
    except OSError:
        return None, "profiles/roofline_traffic.json missing"
    for entry in table:
        if re.search(entry["kernel"], kernel_name) and int(entry.get("n_dim", n_dim)) == n_dim and int(entry.get("n", n)) == n \
                and entry.get("flow", "maf6") == FLOW:
            return float(entry["dram_bytes_per_launch"]), entry["source"]
    return None, f"no committed ncu capture of {kernel_name!r} at n={n}, n_dim={n_dim}, flow={FLOW}"


def tri_issued_flop(tri_meta, n):
    """FLOP the tcgen05 groups of one windowed block-triangular sweep issue, per 128-particle tile and per transform
    (tri_layout.build_tri tables): right-looking updates 2 * 128 * N * K per block and layer, left-looking window
    initialisations 2 * 128 * N * K(all earlier slots) per window and layer; 3 passes (3xTF32)."""
    from pocomc_b200 import tri_layout as TL
    m = np.asarray(tri_meta, np.int64)
    nb, nw = int(m[TL.TRI_NB]), int(m[TL.TRI_NW])
    blocks = m[m[TL.TRI_OFF_BLOCKS]:m[TL.TRI_OFF_BLOCKS] + nb * TL.TB_FIELDS].reshape(nb, -1)
    wins = m[m[TL.TRI_OFF_WINDOWS]:m[TL.TRI_OFF_WINDOWS] + nw * TL.TW_FIELDS].reshape(nw, -1)
    mac = 0.0
    for b in blocks:
        if not (b[TL.TB_FLAGS] & 1):
            mac += b[TL.TB_UPD_N] * (8 + 2 * b[TL.TB_KP]) + b[TL.TB_OUT_N] * b[TL.TB_KP]
    for w in wins[1:]:
        mac += w[TL.TW_WP] * (w[TL.TW_KX] + 2 * w[TL.TW_KH]) + w[TL.TW_OP] * w[TL.TW_KH]
    return 2.0 * 128 * mac * 3 * int(m[TL.TRI_T]) * math.ceil(n / 128)


# ---------------------------------------------------------------------------------------------
# B200 arm
# ---------------------------------------------------------------------------------------------
def run_b200(args):
    import torch
    import pocomc_b200 as pc
    from pocomc_b200 import _lib, config, dist, mcmc as M
    from pocomc_b200 import made_layout as ML
    from pocomc_b200.synthetic import DevicePrior

    rank, world = dist.init_from_env()
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device (the hot path has no CPU fallback); use --impl reference for the CPU arm")
    dev = torch.device("cuda", torch.cuda.current_device())
    if world > 1:       # one process per GPU shares the box's host cores: keep each rank's BLAS / torch pools to its share
        share = max(1, (os.cpu_count() or 1) // world)
        torch.set_num_threads(share)
        try:
            # PMC_BENCH_PIN=1: ... and on its own block of cores (the ranks meet at every MCMC step).  Off by default: on the
            # 8-GPU box of this pool a contiguous split measured no better than the scheduler's own placement (e2e 69.6 M
            # pinned vs 74.7 M unpinned at N = 8; the split ignores which socket a rank's GPU hangs off)
            cores = sorted(os.sched_getaffinity(0)) if os.environ.get("PMC_BENCH_PIN") == "1" else []
            if cores and len(cores) >= world:
                per = len(cores) // world
                mine = set(cores[rank * per:(rank + 1) * per])
                for tid in os.listdir("/proc/self/task"):          # the BLAS / torch worker threads that already exist, too
                    try:
                        os.sched_setaffinity(int(tid), mine)
                    except (OSError, ValueError):
                        pass
        except (AttributeError, OSError):
            pass
        try:
            from threadpoolctl import threadpool_limits
            threadpool_limits(limits=share)
        except ImportError:
            pass
    cfg = CONFIGS[args.config]
    D, MCMC_STEPS = cfg["d"], cfg["steps"]
    if args.scaling == "strong":
        # fixed TOTAL: the config's own total where it is defined on several GPUs, else 64 x the per-GPU count (a total
        # small enough for one tile wave per GPU would only measure the latency floor of one sweep)
        total = cfg["n"] * cfg["gpus"] if cfg["gpus"] > 1 else 64 * cfg["n"]
        # ... rounded down to whole waves of the flow inverse's 128-particle tiles on 148 SMs x 8 GPUs, so that every rank
        # runs full waves at every N in {1, 2, 4, 8} (config 1: 606 208 particles = 4 x 148 tiles per GPU at N = 8; an
        # unrounded 640 000 costs every rank a fifth, 23 %-full wave there)
        wave = 128 * 148 * 8
        total = max(wave, total // wave * wave)
        n_local = (total // world + 255) // 256 * 256
    else:
        n_local = cfg["n"]
    n_global = n_local * world
    wl = Workload(cfg, n_local, row_offset=rank * n_local)

    # ---- untimed setup: scaler, flow training, geometry (identical on every rank) ----
    np.random.seed(0)
    torch.manual_seed(0)
    scaler = pc.scaler.Reparameterize(D, bounds=wl.bounds)
    scaler.fit(wl.prior_samples)
    wl0 = wl if rank == 0 and n_local >= 10_000 else Workload(cfg, min(10_000, max(n_local, 2000)), row_offset=0)   # rank-0 cloud trains the (replicated) flow
    u_train = scaler.forward(wl0.x0[:10_000])
    flow = pc.Flow(D, FLOW)
    flow.fit(torch.tensor(u_train, dtype=torch.float32), validation_split=0.5, epochs=TRAIN_EPOCHS, batch_size=512,
             patience=10 ** 6, annealing=False)
    theta_train = pc.tools.flow_numpy_wrapper(flow).forward(u_train)[0]
    geo = pc.geometry.Geometry()
    geo.fit(theta_train.astype(np.float64))
    u0 = scaler.forward(wl.x0)
    ldj0 = scaler.inverse(u0)[1]
    state = dict(u=u0, x=wl.x0, logdetj=ldj0, logl=wl.loglike(wl.x0), logp=wl.logprior(wl.x0), beta=1.0, blobs=None)
    prior_dev = DevicePrior(np.full(D, wl.prior_kind, np.int32), np.full(D, wl.prior_loc), np.full(D, wl.prior_scale))
    blocks = int(_lib.load().pmc_mh_partials_size(n_local, D)) // (D + 4)
    shard = (rank * n_local, n_global, [blocks] * world) if world > 1 else None
    flush = torch.empty(L2_FLUSH_BYTES, dtype=torch.uint8, device=dev)

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
        return float(t.item())

    # ---- arm 1: device resident -----------------------------------------------------------------
    config.set_rng_mode("device")
    sweep_events = []
    fd = dict(loglike=lambda x: (wl.loglike(x), None), logprior=wl.logprior, scaler=scaler, flow=flow,
              theta_geometry=geo, u_geometry=geo, loglike_device=wl.like.device, logprior_device=prior_dev)
    od = dict(n_max=MCMC_STEPS, n_steps=10 ** 9, progress_bar=None, proposal_scale=2.38 / D ** 0.5, shard=shard,
              seed=1234, sweep_events=None)
    eng = M.McmcEngine(M.KIND_TPCN_FLOW, state, fd, od)

    def device_step():
        flush.fill_(1)
        eng.reset_controller()
        eng.loop()
        assert eng.step == MCMC_STEPS

    for _ in range(args.warmup):
        device_step()
    eng.sweep_events = sweep_events
    clock_lines, stop_evt = [], threading.Event()
    th = threading.Thread(target=clocks_sampler, args=(stop_evt, clock_lines), daemon=True)
    th.start()
    barrier()
    calls0 = _lib.entry_calls()
    torch.cuda.profiler.start()          # ncu --profile-from-start off captures exactly the timed region
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        device_step()
    e1.record()
    barrier()
    torch.cuda.profiler.stop()
    launches = _lib.entry_calls() - calls0          # COUNTED: every libpmc_b200 entry point invoked in the timed region launches >= 1 kernel
    t_dev = max_over_ranks(e0.elapsed_time(e1) * 1e-3)
    sweep_ms = float(np.mean([a.elapsed_time(b) for a, b in sweep_events]))
    eng.sweep_events = None
    accept_dev = eng.accept

    # ---- arm 2: end to end through the reference-facing kernel seam, host buffers ---------------
    prior = pc.Prior(wl.dists)          # what Sampler passes: prior.logpdf (device fast path for norm/uniform factors)
    fd_host = dict(loglike=lambda x: (wl.loglike(x), None), logprior=prior.logpdf, scaler=scaler, flow=flow,
                   theta_geometry=geo, u_geometry=geo)
    od_host = dict(n_max=MCMC_STEPS, n_steps=10 ** 9, progress_bar=None, proposal_scale=2.38 / D ** 0.5, shard=shard, seed=1234)

    def e2e_step():
        flush.fill_(1)
        res = M.preconditioned_pcn(dict(state), fd_host, od_host)      # the call uploads its inputs and never writes to them
        assert res["steps"] == MCMC_STEPS
        return res

    e2e_steps = args.steps if n_local * D <= 2_000_000 else max(1, args.steps // 2)
    for _ in range(max(1, args.warmup // 2)):
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    barrier()
    t_e2e = max_over_ranks(time.perf_counter() - t0)
    stop_evt.set()
    th.join(timeout=3)

    per_call = n_local * (2 * D + 3) * 8
    per_mcmc_d2h = n_local * (D * 8 + 1) + (M.CTL_MU + D) * 8     # x', finite mask, controller block
    per_mcmc_h2d = n_local * 8                                    # logl' (log-prior evaluated on the GPU)
    h2d = per_call + MCMC_STEPS * per_mcmc_h2d
    d2h = per_call + MCMC_STEPS * per_mcmc_d2h

    # ---- roofline of the dominant kernel (flow inverse) -------------------------------------------
    lay = flow.flow.layout
    useful = 2.0 * ML.useful_macs(lay) * n_local                  # 2 * nnz(masks) per particle: the algorithmic minimum
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except OSError:
        pass
    peak_tf = float(peaks.get("bf16_tflops", 1590.0))
    mod = flow.flow
    if mod.tri_available() and config.inverse_path == "tri":
        kernel = "made_sweep_tri_kernel<true> (flow inverse: windowed tcgen05 block-triangular sweep)" if lay.kind == ML.KIND_AFFINE else \
            "made_sweep_tri_kernel<true, rqs> (flow inverse of a spline flow: windowed tcgen05 block-triangular sweep, rational-quadratic head)"
        issued = tri_issued_flop(mod._tri_meta_host, n_local)
        note = ("tcgen05.mma kind::tf32 (3xTF32 split): right-looking updates inside a tensor-memory window, left-looking window "
                "initialisation from the fp32 scratch area, in-block fp32 substitution; achieved = issued MMA FLOP per launch (2*128*N*K per "
                "group, 3 passes, per 128-particle tile and transform) / launch time; one 128-particle tile per SM (TMEM holds one tile's "
                "window) -- the substitution chain, the operand stream from L2 and the tensor pipe take turns (DESIGN.md section 4)")
    else:
        kernel = "made_sweep_stream_kernel<Affine> (flow inverse: fp32-FMA degree-ordered sweep)" if ML.stream_supported(D, lay.n_hidden, lay.n_layers, lay.kind) \
            else "made_sweep_kernel<Affine> (flow inverse: fp32-FMA sweep, weights through L2)"
        issued = useful
        note = "fp32 FMA sweep on CUDA cores (no tensor-core path for this flow shape yet); achieved = useful FLOP (2*nnz(masks)) / launch time"
    achieved_tf = issued / (sweep_ms * 1e-3) / 1e12
    traffic, traffic_src = committed_traffic(kernel, n_local, D)
    roofline = dict(bound="tensor", achieved=achieved_tf, peak=peak_tf, unit="TFLOP/s", frac=achieved_tf / peak_tf,
                    traffic=traffic, traffic_source=traffic_src, kernel=kernel,
                    peak_source="MEASURED_PEAKS.json bf16 burst" if peaks else "fallback 1.59 PFLOP/s",
                    flop_per_launch=issued, useful_flop_per_launch=useful, useful_tflops=useful / (sweep_ms * 1e-3) / 1e12,
                    avg_launch_ms=sweep_ms, share_of_step=sweep_ms * MCMC_STEPS * args.steps / (t_dev * 1e3), note=note)

    # the rest of the device-resident step (rng, proposal, scaler, prior, synthetic likelihood, accept + adapt [+ exchange]): HBM-bound
    hbm_peak = float(peaks.get("hbm_gbs", 6551.0))
    step_ms = 1e3 * t_dev / (args.steps * MCMC_STEPS)
    chain_ms = max(step_ms - sweep_ms, 1e-6)
    chain_bytes = (44.0 * D + 56.0) * n_local                      # algorithmic bytes per particle-step of the non-flow part
    roofline_chain = dict(bound="hbm", kernel="non-flow chain of one MCMC step (rng_fill, tpcn_propose, scaler_inverse, logprior, loglike, mh_accept + adapt)",
                          achieved=chain_bytes / (chain_ms * 1e-3) / 1e9, peak=hbm_peak, unit="GB/s",
                          frac=chain_bytes / (chain_ms * 1e-3) / 1e9 / hbm_peak, bytes_per_step=chain_bytes, ms_per_step=chain_ms,
                          note="step time minus the flow-inverse launch (CUDA events); (44 D + 56) B per particle-step; the state of one shard "
                               "is L2-resident, the chain is launch-latency bound at this size (6 launches)")

    value = n_global * MCMC_STEPS * args.steps / t_dev
    e2e_value = n_global * MCMC_STEPS * e2e_steps / t_e2e
    line = dict(metric="particle-steps/sec", value=value, unit="particle-steps/s", n_gpus=world, steps=args.steps,
                warmup=args.warmup, ms_per_step=1e3 * t_dev / args.steps, higher_is_better=True, scaling=args.scaling,
                vs_baseline=None, dtype="f32 flow / f64 SMC state", data="synthetic",
                config=workload_config(args.config, cfg, n_local, world, MCMC_STEPS, args.scaling), clocks=summarise_clocks(clock_lines),
                e2e=dict(value=e2e_value, unit="particle-steps/s", h2d_bytes_per_step=h2d, d2h_bytes_per_step=d2h,
                         ms_per_step=1e3 * t_e2e / e2e_steps, steps=e2e_steps, rng="device Philox",
                         callbacks="host numpy likelihood (black box); pc.Prior of scipy norm / uniform factors evaluated on the GPU"),
                gpu_launches=launches, roofline=roofline, roofline_other=[roofline_chain], accept_rate=accept_dev)
    if rank == 0 and world == 1 and not args.no_aux and args.config == 1 and FLOW == "maf6":
        line["aux"] = aux_measurements(flow, peaks, D)
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        params = [p.detach().cpu().numpy() for _, p in sorted(flow_param_arrays(flow))]
        run, kind = reference_problem(wl, flow_params=params, threads=threads, n_sub=cfg["cpu_n"])
        dt = run(1)
        if dt < 10.0:
            dt = min(dt, run(1))
        line["cpu_baseline"] = dict(value=run.n / dt, unit="particle-steps/s", cores=threads, kind=kind,
                                    sample=f"1 tpCN step x {run.n} particles ({'the whole shard' if run.n == n_local else 'a sample of the shard'}), same trained "
                                           f"flow weights; {'unmodified reference (oracle/_ref) over oracle/zuko' if kind == 'reference' else 'oracle port'}")
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        torch.distributed.destroy_process_group()


# ---------------------------------------------------------------------------------------------
# auxiliary measurements (rank 0, N = 1, config 1): tensor-core forward, training step, full-run logZ
# ---------------------------------------------------------------------------------------------
def aux_measurements(flow, peaks, n_dim):
    import torch
    import pocomc_b200 as pc
    from pocomc_b200.flow import _FitEngine
    out = {}

    def timeit(fn, reps):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps

    # (0) a MEASURED TF32 GEMM peak (cuBLAS, 8192^3) next to the bf16 one of MEASURED_PEAKS.json
    try:
        a = torch.randn(8192, 8192, device="cuda")
        b = torch.randn(8192, 8192, device="cuda")
        old = torch.backends.cuda.matmul.allow_tf32
        torch.backends.cuda.matmul.allow_tf32 = True
        ms = min(timeit(lambda: torch.matmul(a, b), 5) for _ in range(3))
        torch.backends.cuda.matmul.allow_tf32 = old
        out["tf32_gemm_peak_tflops"] = 2 * 8192 ** 3 / (ms * 1e-3) / 1e12
        del a, b
    except Exception as e:      # diagnostics only
        out["tf32_gemm_peak_tflops"] = repr(e)
    tf32_peak = out["tf32_gemm_peak_tflops"] if isinstance(out["tf32_gemm_peak_tflops"], float) else None
    mod = flow.flow
    # (1) Flow.forward on tcgen05 (csrc/flow_tc.cu): 1 M particles >> L2, 3xTF32
    if mod.tc_available():
        n = 1 << 20
        x = torch.randn(n, n_dim, device="cuda")
        z = torch.empty_like(x)
        l = torch.empty(n, device="cuda")
        lay = mod.layout
        kx, nout, h = (n_dim + 7) // 8 * 8, (2 * n_dim + 15) // 16 * 16, lay.n_hidden
        per_t = 2 * (kx * h + (lay.n_layers - 1) * h * h + h * nout)          # dense masked-MLP pass (biases are added by the epilogue)
        ms3 = timeit(lambda: mod.forward_tc_into(x, z, l, 3), 10)
        issued = 3 * per_t * lay.n_transforms * n / (ms3 * 1e-3) / 1e12
        peak = float(peaks.get("bf16_tflops", 1590.0))
        out["flow_forward_tcgen05"] = dict(particles=n, ms=ms3, particles_per_s=n / (ms3 * 1e-3), issued_tflops=issued,
                                           useful_tflops=issued / 3, frac_of_bf16_peak=issued / peak,
                                           frac_of_measured_tf32_gemm_peak=(issued / tf32_peak) if tf32_peak else None)
        # (1b) the block-triangular inverse at the same size (every SM holds a tile: throughput, not one-wave latency)
        if mod.tri_available():
            ms_inv = timeit(lambda: mod.sweep_tri_into(x, z, l, True), 5)
            iss = tri_issued_flop(mod._tri_meta_host, n) / (ms_inv * 1e-3) / 1e12
            out["flow_inverse_tcgen05"] = dict(particles=n, ms=ms_inv, particles_per_s=n / (ms_inv * 1e-3), issued_tflops=iss,
                                               frac_of_bf16_peak=iss / peak, frac_of_measured_tf32_gemm_peak=(iss / tf32_peak) if tf32_peak else None)
            pc.config.inverse_path = "sweep"
            ms_ffma = timeit(lambda: mod.sweep_into(x[:200_000], z[:200_000], l[:200_000], True), 3) * (n / 200_000)
            pc.config.inverse_path = "tri"
            out["flow_inverse_tcgen05"]["ffma_sweep_ms_extrapolated"] = ms_ffma
        del x, z, l
    # (2) one optimiser step of Flow.fit (batch 512) on the fused kernels inside a CUDA graph
    try:
        f2 = pc.Flow(n_dim, FLOW)
        eng = _FitEngine(f2.flow)
        xt = torch.randn(8192, n_dim, device="cuda")
        wt = torch.rand(8192, device="cuda") + 0.1
        eng.load(xt, wt)
        eng.reset_optimizer(1e-3, 0.0, 1.0)
        batches = [torch.arange(i, i + 512) for i in range(0, 8192, 512)]
        eng.run_epoch(batches, 512, True)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(5):
            eng.run_epoch(batches, 512, True)
        torch.cuda.synchronize()
        us = (time.perf_counter() - t0) / 80 * 1e6
        lay2 = f2.flow.layout
        dense = lay2.n_transforms * (lay2.n_dim * lay2.n_hidden + (lay2.n_layers - 1) * lay2.n_hidden ** 2 + lay2.n_hidden * lay2.n_dim * lay2.total)
        gflop = 6.0 * dense * 512 / 1e9                         # forward + input gradients + weight gradients, dense
        out["fit_step"] = dict(batch=512, flow=FLOW, n_dim=n_dim, us_per_optimizer_step=us, gflop_per_step=gflop,
                               tflops=gflop / us * 1e3, frac_of_fp32_fma_peak=gflop / us * 1e3 / 74.45,
                               path="fused forward/backward + grouped weight-gradient GEMM + clip/AdamW, one CUDA graph launch"
                               if eng.fused else ("layer-wise kernels in a CUDA graph" if eng.layerwise else "autograd in a CUDA graph"))
    except Exception as e:      # diagnostics only
        out["fit_step"] = dict(error=repr(e))
    # (2b) the same for the reference's default flow family (nsf6, 10-D): layer-wise kernels (csrc/flow_train_lw.cu)
    try:
        f3 = pc.Flow(10, "nsf6")
        eng3 = _FitEngine(f3.flow)
        xt = torch.randn(8192, 10, device="cuda")
        wt = torch.rand(8192, device="cuda") + 0.1
        eng3.load(xt, wt)
        eng3.reset_optimizer(1e-3, 0.0, 1.0)
        batches = [torch.arange(i, i + 512) for i in range(0, 8192, 512)]
        eng3.run_epoch(batches, 512, True)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(5):
            eng3.run_epoch(batches, 512, True)
        torch.cuda.synchronize()
        out["fit_step_nsf6_10d"] = dict(batch=512, us_per_optimizer_step=(time.perf_counter() - t0) / 80 * 1e6,
                                        path="layer-wise kernels in a CUDA graph" if eng3.layerwise else "autograd in a CUDA graph")
    except Exception as e:
        out["fit_step_nsf6_10d"] = dict(error=repr(e))
    # (2c) proposal geometry on the device: weighted fit of a 40 000 x n_dim cloud (what Sampler._train hands over at this config)
    try:
        rng = np.random.default_rng(0)
        cloud = rng.normal(size=(40000, n_dim))
        wts = np.exp(rng.normal(size=40000)); wts /= wts.sum()
        geo = pc.geometry.Geometry()
        best = 1e9
        for _ in range(3):
            np.random.seed(1)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            geo.fit(cloud, weights=wts)
            torch.cuda.synchronize()
            best = min(best, time.perf_counter() - t0)
        out["geometry_fit"] = dict(rows=40000, n_dim=n_dim, ms=best * 1e3, path="csrc/geom_ops.cu (f64, fixed-order reductions) + device sort / resample / gather")
    except Exception as e:
        out["geometry_fit"] = dict(error=repr(e))
    # (3) full Sampler.run() of BASELINE configs[0]; logZ against the unmodified reference
    try:
        from scipy.stats import uniform
        gold = json.load(open(os.path.join(ROOT, "tests", "golden", "rosen10.json")))["runs"]

        def rosen(x):
            return -np.sum(10.0 * (x[:, ::2] ** 2.0 - x[:, 1::2]) ** 2.0 + (x[:, ::2] - 1.0) ** 2.0, axis=1)

        runs = []
        for rep in range(2):                                   # the first run pays graph capture / page-in
            smp = pc.Sampler(pc.Prior([uniform(-10.0, 20.0)] * 10), rosen, vectorize=True, n_active=1000, n_effective=2000,
                             flow="maf6", random_state=0)
            t0 = time.perf_counter()
            smp.run(n_total=4096, n_evidence=4096, progress=False)
            wall = time.perf_counter() - t0
            logz, err = smp.evidence()
            steps = int(np.sum(smp.results["steps"]))
            runs.append((wall, float(logz), float(err), steps))
        wall, logz, err, steps = runs[-1]
        ref_logz = float(np.mean([g["logz"] for g in gold]))
        out["full_run"] = dict(workload="BASELINE configs[0]: 10-D Rosenbrock, U(-10,10)^10, n_active=1000, n_effective=2000, maf6, n_total=4096",
                               rng_mode=pc.config.rng_mode, seconds=wall, first_run_seconds=runs[0][0], mcmc_steps=steps, particle_steps_per_s=1000 * steps / wall,
                               logz=logz, logz_err=err, reference_logz=[g["logz"] for g in gold],
                               reference_logz_err=[g["logz_err"] for g in gold], logz_abs_err=abs(logz - ref_logz),
                               reference_cpu_seconds=[g["wall_s"] for g in gold], reference_cpu_cores=gold[0]["cores"],
                               reference_particle_steps_per_s=[g["particle_steps_per_s"] for g in gold],
                               note="reference = unmodified pocomc on the build container's CPU (oracle/make_golden_rosen.py); "
                                    "both logZ carry the sampler's own Monte-Carlo error (logz_err)")
    except Exception as e:
        out["full_run"] = dict(error=repr(e))
    return out


def flow_param_arrays(flow):
    """(index, tensor) per module-order parameter tensor of the flat blob (for the reference / oracle flow)."""
    out, k = [], 0
    for t in range(flow.flow.layout.n_transforms):
        for w, b in flow.flow.transform_params(t):
            out.append((k, w)); out.append((k + 1, b)); k += 2
    return out


def main():
    global FLOW
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", type=int, default=1, choices=sorted(CONFIGS), help="BASELINE configs index (5 = the literal 10k x 32-D Rosenbrock line)")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"])
    ap.add_argument("--flow", default=FLOW, choices=["maf3", "maf6", "maf12", "nsf3", "nsf6", "nsf12"],
                    help="flow preset of both arms (BASELINE's configs use maf6; nsf6 is the reference's own default)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-aux", action="store_true", help="skip the untimed auxiliary measurements (tcgen05 forward, fit step, full run)")
    args = ap.parse_args()
    FLOW = args.flow
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
