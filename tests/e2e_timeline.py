"""Dev helper: where the wall time of the reference-facing preconditioned_pcn call goes (bench.py's e2e arm), phase by phase,
with perf_counter around the engine's own methods (no extra synchronisation: a phase that launches work only pays its launch
cost, the waits show up in the phases that synchronise).   CFG=1 python tests/e2e_timeline.py"""
import os, sys, time, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench as B
import pocomc_b200 as pc
from pocomc_b200 import config, mcmc as M

cfg = B.CONFIGS[int(os.environ.get("CFG", "1"))]
D, N = cfg["d"], cfg["n"]
wl = B.Workload(cfg, N)
np.random.seed(0); torch.manual_seed(0)
scaler = pc.scaler.Reparameterize(D, bounds=wl.bounds); scaler.fit(wl.prior_samples)
u0 = scaler.forward(wl.x0)
flow = pc.Flow(D, B.FLOW)
flow.fit(torch.tensor(u0[:10000], dtype=torch.float32), validation_split=0.5, epochs=3, batch_size=512, patience=10 ** 6, annealing=False)
theta = pc.tools.flow_numpy_wrapper(flow).forward(u0[:10000])[0]
geo = pc.geometry.Geometry(); geo.fit(theta.astype(np.float64))
state = dict(u=u0, x=wl.x0, logdetj=scaler.inverse(u0)[1], logl=wl.loglike(wl.x0), logp=wl.logprior(wl.x0), beta=1.0, blobs=None)
prior = pc.Prior(wl.dists)
acc = collections.defaultdict(float)
cnt = collections.defaultdict(int)


def timed(name, fn):
    def wrap(*a, **k):
        t0 = time.perf_counter()
        out = fn(*a, **k)
        acc[name] += time.perf_counter() - t0
        cnt[name] += 1
        return out
    return wrap


def like(x):
    t0 = time.perf_counter()
    out = wl.loglike(x)
    acc["  likelihood (user function)"] += time.perf_counter() - t0
    return out, None


fd = dict(loglike=like, logprior=prior.logpdf, scaler=scaler, flow=flow, theta_geometry=geo, u_geometry=geo)
od = dict(n_max=50, n_steps=10 ** 9, progress_bar=None, proposal_scale=2.38 / D ** 0.5, seed=1)
config.set_rng_mode("device")
config.host_chunks = int(os.environ.get("CHUNKS", "1"))
M.preconditioned_pcn(dict(state), fd, od)
E = M.McmcEngine
for name in ("__init__", "draw_noise", "propose", "pull_back", "evaluate_host", "accept_and_adapt", "read_controller", "_after_step", "results"):
    setattr(E, name, timed(name, getattr(E, name)))
orig_bind = E._bind_host_io


def bind_io(self):
    orig_bind(self)
    self._download = timed("  download launch", self._download)
    self._events.wait = [timed("  wait for x' chunk", w) for w in self._events.wait]
    self._up_logl = timed("  upload logl launch", self._up_logl)


E._bind_host_io = bind_io
acc.clear(); cnt.clear()
REPS = 5
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(REPS):
    M.preconditioned_pcn(dict(state), fd, od)
tot = time.perf_counter() - t0
steps = REPS * 50
print(f"N={N} D={D} chunks={config.host_chunks}: {1e6 * tot / steps:.1f} us per MCMC step, {N * steps / tot / 1e6:.2f} M particle-steps/s")
for k, v in acc.items():
    per = "per call" if k in ("__init__", "results") else "per step"
    print(f"  {k:32s} {1e6 * v / steps:8.1f} us/step   ({cnt[k]} calls)")
print(f"  {'unaccounted':32s} {1e6 * (tot - sum(v for k, v in acc.items() if not k.startswith('  '))) / steps:8.1f} us/step")
