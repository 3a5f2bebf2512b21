"""tpCN proposal kernel timing (run on the GPU box; not a pytest): python tests/propose_bench.py [n d]..."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pocomc_b200 import _lib
sizes = [(int(sys.argv[i]), int(sys.argv[i + 1])) for i in range(1, len(sys.argv) - 1, 2)] or [(10000, 32), (50000, 50), (50000, 100), (125000, 200)]
dev = torch.device("cuda")
for n, d in sizes:
    rng = np.random.default_rng(0)
    a = rng.normal(size=(d, d)) / np.sqrt(d); cov = a @ a.T + 0.1 * np.eye(d)
    t = lambda arr, dt=torch.float64: torch.as_tensor(np.ascontiguousarray(arr), dtype=dt).to(dev)
    ctl = np.zeros(16 + d); ctl[0] = 0.2
    ctl_d, th, g, z = t(ctl), torch.randn(n, d, device=dev), torch.rand(n, device=dev, dtype=torch.float64) + 1.0, torch.randn(n, d, device=dev, dtype=torch.float64)
    inv_d, chol_d = t(np.linalg.inv(cov).T), t(np.linalg.cholesky(cov).T)
    prop = torch.empty(n, d, dtype=torch.float64, device=dev); p32 = torch.empty(n, d, device=dev)
    mc = torch.empty(n, dtype=torch.float64, device=dev); mp = torch.empty_like(mc)
    rec = dict(n=n, d=d)
    for rows in ("1", "0"):
        os.environ["PMC_TPCN_ROW_KERNEL"] = rows
        run = lambda: _lib.call("pmc_tpcn_propose", 1, _lib.ptr(th), _lib.ptr(ctl_d), _lib.ptr(inv_d), _lib.ptr(chol_d), 5.0, _lib.ptr(g), _lib.ptr(z),
                                _lib.ptr(prop), _lib.ptr(p32), _lib.ptr(mc), _lib.ptr(mp), n, d)
        for _ in range(2): run()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5): run()
        e1.record(); torch.cuda.synchronize()
        rec["row_kernel_us" if rows == "1" else "default_us"] = e0.elapsed_time(e1) / 5 * 1e3
    rec["gdfma_per_s"] = 3.0 * n * d * d / (rec["default_us"] * 1e-6) / 1e9
    print(json.dumps(rec), flush=True)
