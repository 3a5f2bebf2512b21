"""TEST INFRASTRUCTURE -- runs the UNMODIFIED reference Sampler (/root/reference/pocomc, with oracle/zuko
standing in for the absent third-party zuko) on BASELINE configs[0] (10-D Rosenbrock, n_active=1000,
maf6, vectorize=True, README.md:48-75) and records logZ, run length and CPU wall time into
tests/golden/rosen10.json.  Run in the build container only: python oracle/make_golden_rosen.py [seed ...]"""
import json
import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path[:0] = [HERE, "/root/reference"]
import numpy as np  # noqa: E402
import pocomc as pc  # noqa: E402
from scipy.stats import uniform  # noqa: E402

D = 10


def loglike(x):
    return -np.sum(10.0 * (x[:, ::2] ** 2.0 - x[:, 1::2]) ** 2.0 + (x[:, ::2] - 1.0) ** 2.0, axis=1)


def main():
    seeds = [int(a) for a in sys.argv[1:]] or [0]
    out = []
    for seed in seeds:
        prior = pc.Prior([uniform(-10.0, 20.0)] * D)
        s = pc.Sampler(prior, loglike, vectorize=True, n_active=1000, n_effective=2000, flow="maf6", random_state=seed)
        t0 = time.perf_counter()
        s.run(n_total=4096, n_evidence=4096, progress=False)
        wall = time.perf_counter() - t0
        logz, err = s.evidence()
        r = s.results
        out.append(dict(seed=seed, logz=float(logz), logz_err=float(err), iterations=int(len(r["beta"])),
                        mcmc_steps=int(np.sum(r["steps"])), calls=int(s.calls), wall_s=wall, cores=os.cpu_count(),
                        particle_steps_per_s=1000 * int(np.sum(r["steps"])) / wall))
        print(out[-1], flush=True)
    with open(os.path.join(HERE, "..", "tests", "golden", "rosen10.json"), "w") as f:
        json.dump(dict(workload="10-D Rosenbrock, U(-10,10)^10 prior, n_active=1000, n_effective=2000, maf6, n_total=4096, n_evidence=4096",
                       runs=out), f, indent=1)


if __name__ == "__main__":
    main()
