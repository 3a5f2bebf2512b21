"""Geometry.fit on the device against the reference's numpy formulation (run on the GPU box; not a pytest):
python tests/geometry_bench.py [n d]...  prints one JSON line per size (weighted fit, as Sampler._train calls it)."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import pocomc_b200 as pc

sizes = [(int(sys.argv[i]), int(sys.argv[i + 1])) for i in range(1, len(sys.argv) - 1, 2)] or [(40000, 32), (200000, 100), (1000000, 200)]
for n, d in sizes:
    rng = np.random.default_rng(0)
    a = rng.normal(size=(d, d)) / np.sqrt(d)
    x = rng.normal(size=(n, d)) @ a
    w = np.exp(rng.normal(size=n))
    w /= w.sum()
    rec = dict(n=n, d=d)
    for host in (False, True):
        if host and n * d > 5e7:
            rep = 1
        else:
            rep = 3
        geo = pc.geometry.Geometry(host=host)
        best = 1e30
        for _ in range(rep):
            np.random.seed(3)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            geo.fit(x, weights=w)
            torch.cuda.synchronize()
            best = min(best, time.perf_counter() - t0)
        rec["host_s" if host else "device_s"] = best
        if host:
            rec["max_rel_cov_diff"] = float(np.max(np.abs(geo.normal_cov - cov_dev)) / np.max(np.abs(cov_dev)))
            rec["max_rel_tcov_diff"] = float(np.max(np.abs(geo.t_cov - tcov_dev)) / np.max(np.abs(tcov_dev)))
        else:
            cov_dev, tcov_dev = geo.normal_cov.copy(), geo.t_cov.copy()
    print(json.dumps(rec), flush=True)
