"""Dev helper for ncu: a few device-resident MCMC steps of bench.py's config (CFG, default 1) and nothing else timed."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench as B
import pocomc_b200 as pc
from pocomc_b200 import config, mcmc as M
from pocomc_b200.synthetic import DevicePrior

cfg = B.CONFIGS[int(os.environ.get("CFG", "1"))]
D, N = cfg["d"], cfg["n"]
wl = B.Workload(cfg, N)
np.random.seed(0); torch.manual_seed(0)
scaler = pc.scaler.Reparameterize(D, bounds=wl.bounds); scaler.fit(wl.prior_samples)
u0 = scaler.forward(wl.x0)
flow = pc.Flow(D, B.FLOW)
flow.fit(torch.tensor(u0[:10000], dtype=torch.float32), validation_split=0.5, epochs=2, batch_size=512, patience=10 ** 6, annealing=False)
theta = pc.tools.flow_numpy_wrapper(flow).forward(u0[:10000])[0]
geo = pc.geometry.Geometry(); geo.fit(theta.astype(np.float64))
state = dict(u=u0, x=wl.x0, logdetj=scaler.inverse(u0)[1], logl=wl.loglike(wl.x0), logp=wl.logprior(wl.x0), beta=1.0, blobs=None)
prior_dev = DevicePrior(np.full(D, wl.prior_kind, np.int32), np.full(D, wl.prior_loc), np.full(D, wl.prior_scale))
config.set_rng_mode("device")
fd = dict(loglike=lambda x: (wl.loglike(x), None), logprior=wl.logprior, scaler=scaler, flow=flow, theta_geometry=geo, u_geometry=geo,
          loglike_device=wl.like.device, logprior_device=prior_dev)
od = dict(n_max=int(os.environ.get("STEPS", "8")), n_steps=10 ** 9, progress_bar=None, proposal_scale=2.38 / D ** 0.5, seed=1234)
eng = M.McmcEngine(M.KIND_TPCN_FLOW, state, fd, od)
eng.loop()
torch.cuda.synchronize()
print("steps", eng.step, "accept", eng.accept)
