"""Dev helper: cProfile of the reference-facing preconditioned_pcn call (bench.py's e2e arm)."""
import cProfile, pstats, sys, os, io
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench as B
import pocomc_b200 as pc
from pocomc_b200 import config, mcmc as M

cfg = B.CONFIGS[int(os.environ.get("CFG", "1"))]
N_DIM = cfg["d"]
wl = B.Workload(cfg, cfg["n"])
np.random.seed(0); torch.manual_seed(0)
scaler = pc.scaler.Reparameterize(N_DIM, bounds=wl.bounds); scaler.fit(wl.prior_samples)
u0 = scaler.forward(wl.x0)
flow = pc.Flow(N_DIM, B.FLOW)
flow.fit(torch.tensor(u0, dtype=torch.float32), validation_split=0.5, epochs=3, batch_size=512, patience=10**6, annealing=False)
theta = pc.tools.flow_numpy_wrapper(flow).forward(u0)[0]
geo = pc.geometry.Geometry(); geo.fit(theta.astype(np.float64))
state = dict(u=u0, x=wl.x0, logdetj=scaler.inverse(u0)[1], logl=wl.loglike(wl.x0), logp=wl.logprior(wl.x0), beta=1.0, blobs=None)
prior = pc.Prior(wl.dists)
fd = dict(loglike=lambda x: (wl.loglike(x), None), logprior=prior.logpdf, scaler=scaler, flow=flow, theta_geometry=geo, u_geometry=geo)
od = dict(n_max=50, n_steps=10**9, progress_bar=None, proposal_scale=2.38 / N_DIM ** 0.5, seed=1)
config.set_rng_mode("device")
M.preconditioned_pcn(dict(state), fd, od)
pr = cProfile.Profile(); pr.enable()
for _ in range(3):
    M.preconditioned_pcn(dict(state), fd, od)
pr.disable()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(22); print(s.getvalue()[:4500])
