"""The one number the reference publishes that was produced with REAL zuko: docs/source/quickstart.ipynb:233 --
10-D Rosenbrock, N(0, 3^2) prior, default Sampler (nsf6, n_active 256, n_effective 512), random_state 0:
logZ = -21.4303 +- 0.0267 after 39 iterations.  Not a pytest: run as

    python tests/quickstart_pin.py gpu 0 1 2 3 4        # pocomc_b200 on the GPU box
    python tests/quickstart_pin.py ref 0 1 2 3          # unmodified reference over oracle/zuko (CPU, ~4.5 min per seed)

and prints one JSON line per seed.  The statistic (mean over seeds within 2 sigma of the published value, and the
published value inside the seed scatter) is the pin DESIGN.md section 2 quotes for oracle/zuko and for the GPU path."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
mode, seeds = sys.argv[1], [int(s) for s in sys.argv[2:]] or [0]
if mode == "ref":
    ref = os.path.join(ROOT, "oracle", "_ref")
    sys.path[:0] = [os.path.join(ROOT, "oracle"), ref if os.path.isdir(ref) else "/root/reference"]
    import pocomc as pc
else:
    sys.path.insert(0, ROOT)
    import pocomc_b200 as pc
import numpy as np
from scipy.stats import norm

n_dim = 10


def log_likelihood(x):
    return -np.sum(10.0 * (x[:, ::2] ** 2.0 - x[:, 1::2]) ** 2.0 + (x[:, ::2] - 1.0) ** 2.0, axis=1)


for seed in seeds:
    np.random.seed(0)
    prior = pc.Prior(n_dim * [norm(0.0, 3.0)])
    t0 = time.time()
    s = pc.Sampler(prior=prior, likelihood=log_likelihood, vectorize=True, random_state=seed)
    s.run(progress=False)
    logz, err = s.evidence()
    print(json.dumps(dict(mode=mode, seed=seed, iterations=int(s.t), calls=int(s.calls), logz=float(logz), err=float(err),
                          seconds=round(time.time() - t0, 1), published=dict(logz=-21.43033860019544, err=0.026709458540766708,
                                                                              iterations=39, calls=51456))), flush=True)
