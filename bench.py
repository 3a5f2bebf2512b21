#!/usr/bin/env python
"""bench.py -- particle-steps/s of pocoMC's flow-preconditioned MCMC hot path (BASELINE.json).

Workload (configs[1]): 32-D correlated Gaussian likelihood
(C = 0.95 11^T + 0.05 I), prior N(0, 3^2)^32, n_active = 10 000 particles per GPU, flow 'maf6'
(H = 128) trained for 200 optimiser steps, Student-t geometry fitted on the latent cloud, beta = 1.
One bench "step" = one `_mutate`-sized call of the t-preconditioned Crank-Nicolson kernel
(pocomc/mcmc.py:8-183): MCMC_STEPS Metropolis steps over every particle with the plateau rule
disabled.  particle-steps/s = particles x MCMC steps / time.

  value : device-resident arm -- state in HBM, Philox noise and the synthetic prior/likelihood
          evaluated on the GPU, no host traffic inside the timed region.
  e2e   : the reference-facing call `pocomc_b200.mcmc.preconditioned_pcn(state_dict, function_dict,
          option_dict)` with HOST numpy buffers, the likelihood and the scipy prior as host black
          boxes (x' D2H and logl'/logp' H2D every MCMC step), state H2D and result D2H per call.
  aux   : untimed-region extras on rank 0 at N=1 -- the tcgen05 dense flow forward (issued TFLOP/s against the
          measured tensor peak), one Flow.fit optimiser step on the fused training kernels, and a FULL
          Sampler.run() of BASELINE configs[0] (10-D Rosenbrock, 1000 particles) whose logZ is compared with
          the unmodified reference's (tests/golden/rosen10.json) -- the "logZ abs-err vs ref" half of the metric.
  --impl reference : the reference's own CPU path for the same call.  pocoMC is pure Python and
          needs the third-party zuko (absent here and on the GPU box), so the arm runs the in-repo
          CPU oracle port (oracle/smc_ref.py + oracle/zuko) on all host threads.
"""
from __future__ import annotations

import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_PER_GPU = 10_000
N_DIM = 32
FLOW = "maf6"
MCMC_STEPS = 50          # Metropolis steps per bench step (n_max of SURVEY 8d-2)
REF_MCMC_STEPS = 1       # bounded sample for the CPU arms (one step is ~seconds on a CPU)
TRAIN_EPOCHS = 20        # x 10 batches of 512 (train split = 5000 rows) = 200 optimiser steps
L2_FLUSH_BYTES = 256 << 20


# ---------------------------------------------------------------------------------------------
# workload (numpy / scipy only: shared by every arm)
# ---------------------------------------------------------------------------------------------
class Workload:
    def __init__(self, n, d, seed=0, row_offset=0):
        from scipy.stats import norm
        self.n, self.d = n, d
        self.cov = 0.95 * np.ones((d, d)) + 0.05 * np.eye(d)
        self.prec = np.linalg.inv(self.cov)
        self.c0 = -0.5 * (d * math.log(2 * math.pi) + np.linalg.slogdet(self.cov)[1])
        self.prior_sd = 3.0
        self.dists = [norm(0.0, self.prior_sd)] * d
        rng = np.random.default_rng(seed)
        self.prior_samples = rng.normal(0.0, self.prior_sd, size=(2 * N_PER_GPU, d))     # scaler.fit input (same on all ranks)
        chol = np.linalg.cholesky(self.cov)
        rng_x = np.random.default_rng([seed, 1, row_offset])
        self.x0 = rng_x.normal(size=(n, d)) @ chol.T
        self.bounds = np.array([dd.support() for dd in self.dists])

    def loglike(self, x):
        return -0.5 * np.einsum("ij,ij->i", x @ self.prec, x) + self.c0

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


# ---------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the oracle port on host cores
# ---------------------------------------------------------------------------------------------
def cpu_problem(wl, flow_params=None, threads=None):
    """Build the CPU-side problem with the oracle (test infrastructure used as the timed CPU baseline)."""
    import torch
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import flow_ref as F
    import smc_ref as O
    if threads:
        torch.set_num_threads(threads)
    scaler = O.scaler_fit(wl.prior_samples, wl.bounds[:, 0], wl.bounds[:, 1])
    u0 = O.scaler_forward(wl.x0, scaler)
    _, ldj0 = O.scaler_inverse(u0, scaler)
    torch.manual_seed(0)
    flow = F.make_flow(wl.d, FLOW)
    if flow_params is not None:
        F.load_params(flow, flow_params)
    else:
        F.fit(flow, torch.tensor(u0, dtype=torch.float32), validation_split=0.5, epochs=TRAIN_EPOCHS, batch_size=512,
              patience=10 ** 6)
    nf = F.NumpyFlow(flow)
    theta0, _ = nf.forward(u0)
    t_mean, t_cov, t_nu = O.fit_mvstud(theta0.astype(np.float64))
    if not np.isfinite(t_nu):
        t_nu = 1e6
    geo = dict(t_mean=t_mean, t_cov=t_cov, t_nu=t_nu)
    state = dict(u=u0, x=wl.x0, logdetj=ldj0, logl=wl.loglike(wl.x0), logp=wl.logprior(wl.x0), beta=1.0)
    return O, nf, scaler, geo, state


def cpu_steps(O, nf, scaler, geo, state, wl, mcmc_steps):
    t0 = time.perf_counter()
    res = O.mcmc_kernel("tpcn_flow", state, wl.loglike, wl.logprior, scaler, geo,
                        dict(n_max=mcmc_steps, n_steps=10 ** 9, proposal_scale=2.38 / wl.d ** 0.5), flow=nf)
    dt = time.perf_counter() - t0
    assert res["steps"] == mcmc_steps
    return dt


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    wl = Workload(N_PER_GPU, N_DIM)
    np.random.seed(0)
    prob = cpu_problem(wl, threads=threads)
    for _ in range(args.warmup):
        cpu_steps(*prob, wl, REF_MCMC_STEPS)
    times = [cpu_steps(*prob, wl, REF_MCMC_STEPS) for _ in range(args.steps)]
    total = float(np.sum(times))
    value = wl.n * REF_MCMC_STEPS * args.steps / total
    sample = f"{REF_MCMC_STEPS} tpCN step(s) x {wl.n} particles per bench step (oracle port of mcmc.py:8-183 + zuko MAF inverse)"
    line = dict(impl="reference", metric="particle-steps/sec", value=value, unit="particle-steps/s", n_gpus=args.gpus,
                steps=args.steps, warmup=args.warmup, ms_per_step=1e3 * total / args.steps, higher_is_better=True,
                scaling="weak", vs_baseline=None, dtype="f32 flow / f64 SMC state", data="synthetic",
                config=workload_config(1, REF_MCMC_STEPS),
                cpu_baseline=dict(value=value, unit="particle-steps/s", cores=threads, kind="port", sample=sample),
                e2e=dict(value=value, unit="particle-steps/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0))
    print(json.dumps(line))


def workload_config(n_gpus, mcmc_steps):
    return {"workload": f"{N_DIM}-D correlated Gaussian, n_particles={N_PER_GPU} per GPU, flow-precond MCMC only "
                        f"(tpCN, {FLOW}, beta=1)", "n_particles_per_gpu": N_PER_GPU, "n_particles_total": N_PER_GPU * n_gpus,
            "n_dim": N_DIM, "flow": FLOW, "mcmc_steps_per_bench_step": mcmc_steps, "kernel": "preconditioned_pcn",
            "parallelism": f"particle-shard x{n_gpus}" if n_gpus > 1 else "single GPU",
            "l2": "state (~10 MB) is L2-resident within a bench step by design; L2 flushed (256 MiB write) between bench steps"}


# ---------------------------------------------------------------------------------------------
# B200 arm
# ---------------------------------------------------------------------------------------------
def run_b200(args):
    import torch
    import pocomc_b200 as pc
    from pocomc_b200 import config, dist, mcmc as M
    from pocomc_b200 import made_layout as ML
    from pocomc_b200.synthetic import CorrelatedGaussian, DevicePrior

    rank, world = dist.init_from_env()
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device (the hot path has no CPU fallback); use --impl reference for the CPU arm")
    dev = torch.device("cuda", torch.cuda.current_device())
    if world > 1:       # one process per GPU shares the box's host cores: keep each rank's BLAS / torch pools to its share
        share = max(1, (os.cpu_count() or 1) // world)
        torch.set_num_threads(share)
        try:
            from threadpoolctl import threadpool_limits
            threadpool_limits(limits=share)
        except ImportError:
            pass
    n_local = N_PER_GPU
    n_global = n_local * world
    wl = Workload(n_local, N_DIM, row_offset=rank * n_local)

    # ---- untimed setup: scaler, flow training, geometry (identical on every rank) ----
    np.random.seed(0)
    torch.manual_seed(0)
    scaler = pc.scaler.Reparameterize(N_DIM, bounds=wl.bounds)
    scaler.fit(wl.prior_samples)
    wl0 = Workload(N_PER_GPU, N_DIM, row_offset=0)            # rank-0 cloud trains the (replicated) flow
    u_train = scaler.forward(wl0.x0)
    flow = pc.Flow(N_DIM, FLOW)
    flow.fit(torch.tensor(u_train, dtype=torch.float32), validation_split=0.5, epochs=TRAIN_EPOCHS, batch_size=512,
             patience=10 ** 6, annealing=False)
    theta_train = pc.tools.flow_numpy_wrapper(flow).forward(u_train)[0]
    geo = pc.geometry.Geometry()
    geo.fit(theta_train.astype(np.float64))
    u0 = scaler.forward(wl.x0)
    ldj0 = scaler.inverse(u0)[1]
    state = dict(u=u0, x=wl.x0, logdetj=ldj0, logl=wl.loglike(wl.x0), logp=wl.logprior(wl.x0), beta=1.0, blobs=None)
    like_dev = CorrelatedGaussian(N_DIM)
    prior_dev = DevicePrior(np.zeros(N_DIM, np.int32), np.zeros(N_DIM), np.full(N_DIM, wl.prior_sd))
    blocks = int(pc._lib.load().pmc_mh_partials_size(n_local, N_DIM)) // (N_DIM + 4)
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
              theta_geometry=geo, u_geometry=geo, loglike_device=like_dev.device, logprior_device=prior_dev)
    od = dict(n_max=MCMC_STEPS, n_steps=10 ** 9, progress_bar=None, proposal_scale=2.38 / N_DIM ** 0.5, shard=shard,
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
    torch.cuda.profiler.start()          # ncu --profile-from-start off captures exactly the timed region
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        device_step()
    e1.record()
    barrier()
    torch.cuda.profiler.stop()
    t_dev = max_over_ranks(e0.elapsed_time(e1) * 1e-3)
    launches = eng.launches
    eng.launches = 0
    sweep_ms = float(np.mean([a.elapsed_time(b) for a, b in sweep_events]))
    eng.sweep_events = None
    accept_dev = eng.accept

    # ---- arm 2: end to end through the reference-facing kernel seam, host buffers ---------------
    prior = pc.Prior(wl.dists)          # what Sampler passes: prior.logpdf (device fast path for norm/uniform factors)
    fd_host = dict(loglike=lambda x: (wl.loglike(x), None), logprior=prior.logpdf, scaler=scaler, flow=flow,
                   theta_geometry=geo, u_geometry=geo)
    od_host = dict(n_max=MCMC_STEPS, n_steps=10 ** 9, progress_bar=None, proposal_scale=2.38 / N_DIM ** 0.5, shard=shard, seed=1234)

    def e2e_step():
        flush.fill_(1)
        res = M.preconditioned_pcn({k: (v.copy() if isinstance(v, np.ndarray) else v) for k, v in state.items()}, fd_host, od_host)
        assert res["steps"] == MCMC_STEPS
        return res

    for _ in range(max(1, args.warmup // 2)):
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        e2e_step()
    barrier()
    t_e2e = max_over_ranks(time.perf_counter() - t0)
    stop_evt.set()
    th.join(timeout=3)

    per_call_h2d = n_local * (2 * N_DIM + 3) * 8
    per_call_d2h = n_local * (2 * N_DIM + 3) * 8
    per_mcmc_d2h = n_local * (N_DIM * 8 + 1) + (M.CTL_MU + N_DIM) * 8     # x', finite mask, controller block
    per_mcmc_h2d = n_local * 8                                             # logl' (log-prior evaluated on the GPU)
    h2d = per_call_h2d + MCMC_STEPS * per_mcmc_h2d
    d2h = per_call_d2h + MCMC_STEPS * per_mcmc_d2h

    # ---- roofline of the dominant kernel (flow inverse sweep) ------------------------------------
    lay = flow.flow.layout
    macs = ML.useful_macs(lay)                      # per particle, all transforms
    flop = 2.0 * macs * n_local
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except OSError:
        pass
    peak_tf = float(peaks.get("bf16_tflops", 1590.0))
    achieved_tf = flop / (sweep_ms * 1e-3) / 1e12
    roofline = dict(bound="tensor", achieved=achieved_tf, peak=peak_tf, unit="TFLOP/s", frac=achieved_tf / peak_tf,
                    traffic=2084608.0,      # bytes per launch: dram read 2.08 MB + write 0 (ncu --set full, profiles/r1c_sweep_v3_ncu.txt)
                    kernel="made_sweep_stream_kernel<Affine> (flow inverse, degree-ordered sweep)",
                    peak_source="MEASURED_PEAKS.json bf16 burst" if peaks else "fallback 1.59 PFLOP/s",
                    flop_per_launch=flop, avg_launch_ms=sweep_ms,
                    note="fp32 FMA sweep on CUDA cores, bound by shared-memory wavefronts (128 FMA per 5 wavefronts caps the FMA pipe at 20 %, "
                         "DESIGN.md section 7); 2*nnz(masks) useful FLOP per particle; share of step = "
                         f"{sweep_ms * MCMC_STEPS * args.steps / (t_dev * 1e3):.2f}")

    value = n_global * MCMC_STEPS * args.steps / t_dev
    e2e_value = n_global * MCMC_STEPS * args.steps / t_e2e
    line = dict(metric="particle-steps/sec", value=value, unit="particle-steps/s", n_gpus=world, steps=args.steps,
                warmup=args.warmup, ms_per_step=1e3 * t_dev / args.steps, higher_is_better=True, scaling="weak",
                vs_baseline=None, dtype="f32 flow / f64 SMC state", data="synthetic",
                config=workload_config(world, MCMC_STEPS), clocks=summarise_clocks(clock_lines),
                e2e=dict(value=e2e_value, unit="particle-steps/s", h2d_bytes_per_step=h2d, d2h_bytes_per_step=d2h,
                         ms_per_step=1e3 * t_e2e / args.steps, rng="device Philox", callbacks="host numpy likelihood (black box); pc.Prior of scipy norm factors evaluated on the GPU"),
                gpu_launches=launches, roofline=roofline, accept_rate=accept_dev)
    line["roofline"].update(fp32_fma_peak_tflops=148 * 128 * 2 * 1.965e9 / 1e12,
                            frac_of_fp32_fma_peak=achieved_tf / (148 * 128 * 2 * 1.965e9 / 1e12))
    if rank == 0 and world == 1 and not args.no_aux:
        line["aux"] = aux_measurements(flow, peaks)
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        params = [p.detach().cpu().numpy() for _, p in sorted(flow_param_arrays(flow))]
        prob = cpu_problem(wl, flow_params=params, threads=threads)
        dt = cpu_steps(*prob, wl, REF_MCMC_STEPS)
        dt = min(dt, cpu_steps(*prob, wl, REF_MCMC_STEPS))
        line["cpu_baseline"] = dict(value=wl.n * REF_MCMC_STEPS / dt, unit="particle-steps/s", cores=threads, kind="port",
                                    sample=f"{REF_MCMC_STEPS} tpCN step x {wl.n} particles, same trained flow weights, best of 2 "
                                           "(oracle port: vectorised numpy + zuko-restated MAF inverse with D+1 passes)")
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        torch.distributed.destroy_process_group()



# ---------------------------------------------------------------------------------------------
# auxiliary measurements (rank 0, N = 1): tensor-core forward, training step, full-run logZ
# ---------------------------------------------------------------------------------------------
def aux_measurements(flow, peaks):
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

    # (1) Flow.forward on tcgen05 (csrc/flow_tc.cu): 1 M particles >> L2, 3xTF32
    mod = flow.flow
    if mod.tc_available():
        n = 1 << 20
        x = torch.randn(n, N_DIM, device="cuda")
        z = torch.empty_like(x)
        l = torch.empty(n, device="cuda")
        lay = mod.layout
        kx, nout, h = (N_DIM + 7) // 8 * 8, (2 * N_DIM + 15) // 16 * 16, lay.n_hidden
        # issued MMA FLOP per particle: per transform 3 passes x 2 x (Kx*H + (L-1)*H*H + H*Nout) + the bias k-steps
        per_t = 2 * (kx * h + (lay.n_layers - 1) * h * h + h * nout) + 2 * 8 * (lay.n_layers * h + nout)
        ms3 = timeit(lambda: mod.forward_tc_into(x, z, l, 3), 10)
        ms_sweep = timeit(lambda: mod.sweep_into(x[:100_000], z[:100_000], l[:100_000], False), 3) * (n / 100_000)
        issued = 3 * per_t * lay.n_transforms * n / (ms3 * 1e-3) / 1e12
        peak = float(peaks.get("bf16_tflops", 1590.0))
        out["flow_forward_tcgen05"] = dict(particles=n, ms=ms3, particles_per_s=n / (ms3 * 1e-3), issued_tflops=issued,
                                           useful_tflops=issued / 3, frac_of_bf16_peak=issued / peak,
                                           frac_of_tf32_peak=issued / (peak / 2), sweep_kernel_ms_extrapolated=ms_sweep,
                                           note="kind::tf32 MMAs, 3-pass split for fp32 fidelity; TF32 peak taken as half the measured bf16 peak")
        del x, z, l
    # (2) one optimiser step of Flow.fit (batch 512) on the fused kernels inside a CUDA graph
    try:
        f2 = pc.Flow(N_DIM, FLOW)
        eng = _FitEngine(f2.flow)
        xt = torch.randn(8192, N_DIM, device="cuda")
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
        out["fit_step"] = dict(batch=512, flow=FLOW, n_dim=N_DIM, us_per_optimizer_step=(time.perf_counter() - t0) / 80 * 1e6,
                               path="fused forward/backward + grouped weight-gradient GEMM + clip/AdamW, one CUDA graph launch"
                               if eng.fused else "autograd in a CUDA graph")
    except Exception as e:      # diagnostics only
        out["fit_step"] = dict(error=repr(e))
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
    """(index, tensor) per module-order parameter tensor of the flat blob (for the oracle flow)."""
    out, k = [], 0
    for t in range(flow.flow.layout.n_transforms):
        for w, b in flow.flow.transform_params(t):
            out.append((k, w)); out.append((k + 1, b)); k += 2
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-aux", action="store_true", help="skip the untimed auxiliary measurements (tcgen05 forward, fit step, full run)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
