"""Phase breakdown of a full Sampler.run() (run on the GPU box; not a pytest).
usage: python tests/run_profile.py [rosen10|mix50] [n_active]"""
import sys, os, time, json, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from scipy.stats import uniform
import pocomc_b200 as pc

which = sys.argv[1] if len(sys.argv) > 1 else "rosen10"
n_active = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
if which == "rosen10":
    D = 10
    def loglike(x):
        return -np.sum(10.0 * (x[:, ::2] ** 2.0 - x[:, 1::2]) ** 2.0 + (x[:, ::2] - 1.0) ** 2.0, axis=1)
    prior = pc.Prior([uniform(-10.0, 20.0)] * D)
    truth = None
elif which == "gauss32":
    D = 32
    from scipy.stats import norm
    cov = 0.95 * np.ones((D, D)) + 0.05 * np.eye(D)
    prec = np.linalg.inv(cov)
    c0 = -0.5 * (D * np.log(2 * np.pi) + np.linalg.slogdet(cov)[1])
    def loglike(x):
        return -0.5 * np.sum((x @ prec) * x, axis=1) + c0
    prior = pc.Prior([norm(0.0, 3.0)] * D)
    truth = -0.5 * (D * np.log(2 * np.pi) + np.linalg.slogdet(cov + 9.0 * np.eye(D))[1])
else:
    D = 50
    c, sg = 1.5, 0.5
    def loglike(x):
        a = -0.5 * np.sum(((x - c) / sg) ** 2, axis=1); b = -0.5 * np.sum(((x + c) / sg) ** 2, axis=1)
        return np.logaddexp(a, b) - np.log(2.0) - D * (np.log(sg) + 0.5 * np.log(2 * np.pi))
    prior = pc.Prior([uniform(-10.0, 20.0)] * D)
    truth = -D * np.log(20.0)

times = collections.defaultdict(float)
counts = collections.Counter()
def wrap(obj, name):
    fn = getattr(obj, name)
    def inner(*a, **k):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        out = fn(*a, **k)
        torch.cuda.synchronize(); times[name] += time.perf_counter() - t0; counts[name] += 1
        return out
    setattr(obj, name, inner)

s = pc.Sampler(prior, loglike, vectorize=True, n_active=n_active, n_effective=2 * n_active, flow="maf6", random_state=0)
for nm in ("_reweight", "_train", "_resample", "_mutate", "_compute_evidence", "_not_termination"):
    wrap(s, nm)
_fit = s.flow.fit
epochs = [0]
def fit_counted(*a, **k):
    h = _fit(*a, **k)
    epochs[0] += len(h["loss"])
    return h
s.flow.fit = fit_counted
wrap(s.flow, "fit")
t0 = time.perf_counter()
s.run(n_total=max(4096, n_active), n_evidence=4096, progress=False)
total = time.perf_counter() - t0
logz, err = s.evidence()
res = s.results
steps = int(np.sum(res["steps"]))
out = dict(workload=which, n_active=n_active, D=D, total_s=total, logz=float(logz), logz_err=float(err) if err is not None else None,
           truth=truth, iters=int(len(res["beta"])), mcmc_steps=steps, particle_steps_per_s=n_active * steps / total,
           phases={k: round(v, 3) for k, v in times.items()}, counts=dict(counts), fit_epochs=epochs[0],
           fit_graph_replays=getattr(s.flow.flow.__dict__.get("_fit_engine"), "launches", None))
print(json.dumps(out))
