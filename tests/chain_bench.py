"""Dev helper: live CUDA-event times of the non-flow kernels of one MCMC step (bench.py's config, warm L2), bit-identity of
the fused / tiled variants against the plain ones, and the end-to-end call with 1 vs 4 download chunks.

    CFG=1 PMC_TPCN_TILED_MIN_D=32 python tests/chain_bench.py
"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench as B
import pocomc_b200 as pc
from pocomc_b200 import config, mcmc as M
from pocomc_b200.synthetic import DevicePrior

cfg = B.CONFIGS[int(os.environ.get("CFG", "1"))]
D = cfg["d"]
wl = B.Workload(cfg, cfg["n"])
N = cfg["n"]
np.random.seed(0); torch.manual_seed(0)
scaler = pc.scaler.Reparameterize(D, bounds=wl.bounds); scaler.fit(wl.prior_samples)
u0 = scaler.forward(wl.x0)
flow = pc.Flow(D, B.FLOW)
flow.fit(torch.tensor(u0[:10000], dtype=torch.float32), validation_split=0.5, epochs=3, batch_size=512, patience=10 ** 6, annealing=False)
theta = pc.tools.flow_numpy_wrapper(flow).forward(u0[:10000])[0]
geo = pc.geometry.Geometry(); geo.fit(theta.astype(np.float64))
state = dict(u=u0, x=wl.x0, logdetj=scaler.inverse(u0)[1], logl=wl.loglike(wl.x0), logp=wl.logprior(wl.x0), beta=1.0, blobs=None)
prior_dev = DevicePrior(np.full(D, wl.prior_kind, np.int32), np.full(D, wl.prior_loc), np.full(D, wl.prior_scale))
config.set_rng_mode("device")
STEPS = 50


def engine(fused, n_max=STEPS):
    config.fuse_prior = fused
    fd = dict(loglike=lambda x: (wl.loglike(x), None), logprior=wl.logprior, scaler=scaler, flow=flow, theta_geometry=geo, u_geometry=geo,
              loglike_device=wl.like.device, logprior_device=prior_dev)
    od = dict(n_max=n_max, n_steps=10 ** 9, progress_bar=None, proposal_scale=2.38 / D ** 0.5, seed=1234)
    return M.McmcEngine(M.KIND_TPCN_FLOW, state, fd, od)


def timeit(fn, reps=200):
    for _ in range(10):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return 1e3 * e0.elapsed_time(e1) / reps


print(f"config {cfg['name'] if 'name' in cfg else os.environ.get('CFG', '1')}: N={N} D={D} tiled_min_d={os.environ.get('PMC_TPCN_TILED_MIN_D', '64')}")
# ---- bit identity: fused scaler + prior vs the two launches ------------------------------------------------------
ea, eb = engine(False), engine(True)
for e in (ea, eb):
    e.draw_noise(); e.propose(); e.pull_back(); e.evaluate_device()
torch.cuda.synchronize()
same = all(torch.equal(getattr(ea, k), getattr(eb, k)) or
           bool(((getattr(ea, k) == getattr(eb, k)) | (torch.isnan(getattr(ea, k).double()) & torch.isnan(getattr(eb, k).double()))).all())
           for k in ("u_p", "x_p", "ldj_p", "finite", "logp_p", "logl_p"))
print("fused scaler+prior bit-identical to scaler, prior:", same)
# ---- bit identity: tiled vs row proposal -------------------------------------------------------------------------
os.environ["PMC_TPCN_ROW_KERNEL"] = "1"
ea.propose(); torch.cuda.synchronize()
ref = [t.clone() for t in (ea.prop64, ea.prop32, ea.m_cur, ea.m_prop)]
os.environ["PMC_TPCN_ROW_KERNEL"] = "0"
ea.propose(); torch.cuda.synchronize()
print("tiled proposal bit-identical to row proposal:", all(torch.equal(a, b) for a, b in zip(ref, (ea.prop64, ea.prop32, ea.m_cur, ea.m_prop))))

# ---- stage times -------------------------------------------------------------------------------------------------
eng = engine(True, n_max=10 ** 8)
eng.draw_noise(); eng.propose(); eng.pull_back(); eng.evaluate_device()
print("rng_fill            %7.2f us" % timeit(eng.draw_noise))
os.environ["PMC_TPCN_ROW_KERNEL"] = "1"
print("propose (row)       %7.2f us" % timeit(eng.propose))
os.environ["PMC_TPCN_ROW_KERNEL"] = "0"
print("propose (default)   %7.2f us" % timeit(eng.propose))
print("flow inverse        %7.2f us" % timeit(eng._sweep))
print("scaler+prior fused  %7.2f us" % timeit(eb._scaler_inverse))
print("scaler              %7.2f us" % timeit(ea._scaler_inverse))
print("prior               %7.2f us" % timeit(lambda: ea.logprior_device(ea.x_p, ea.finite, ea.logp_p)))
print("loglike (synthetic) %7.2f us" % timeit(lambda: eng.loglike_device(eng.x_p, eng.finite, eng.logl_p)))


def accept():
    eng.accept_and_adapt(None)


print("accept + adapt      %7.2f us" % timeit(accept))


def run_steps(e):
    e.reset_controller()
    e.loop()


for name, e in (("unfused prior", ea), ("fused prior", eb)):
    e.ctl[M.CTL_STOP] = 0.0
    t = timeit(lambda: run_steps(e), reps=5)
    print(f"device-resident step ({name}): %7.2f us per MCMC step, %.1f M particle-steps/s" % (t / STEPS, N * STEPS / t))

# ---- end to end ----------------------------------------------------------------------------------------------------
prior = pc.Prior(wl.dists)
fd_host = dict(loglike=lambda x: (wl.loglike(x), None), logprior=prior.logpdf, scaler=scaler, flow=flow, theta_geometry=geo, u_geometry=geo)
od_host = dict(n_max=STEPS, n_steps=10 ** 9, progress_bar=None, proposal_scale=2.38 / D ** 0.5, seed=1234)
res = {}
for chunks in (1, 2, 4, 8, 0):
    config.host_chunks = chunks
    out = M.preconditioned_pcn(dict(state), fd_host, od_host)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(5):
        out = M.preconditioned_pcn(dict(state), fd_host, od_host)
    dt = (time.perf_counter() - t0) / 5
    res[chunks] = out
    print(f"e2e host_chunks={chunks}: %7.1f us per MCMC step, %.2f M particle-steps/s (accept %.4f)" % (1e6 * dt / STEPS, N * STEPS / dt / 1e6, out["accept"]))
print("e2e results identical across chunkings:", all(np.array_equal(res[1][k], res[c][k]) for c in res for k in ("x", "u", "logl", "logp")))
