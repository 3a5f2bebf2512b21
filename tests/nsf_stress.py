"""Dev helper: spline flows, tcgen05 block-triangular sweep vs the fp32-FMA sweep (reference-ordered spline head) on inputs that
reach past the spline's bound (|x| > 5: identity tails), sit on it, and on trained-scale weights; both directions."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import pocomc_b200 as pc
from pocomc_b200 import config
for d, preset, scale, wmul in ((10, "nsf6", 3.0, 1.0), (10, "nsf6", 1.0, 2.0), (32, "nsf6", 2.5, 1.3), (12, "nsf3", 6.0, 1.5), (50, "nsf3", 2.0, 1.2)):
    torch.manual_seed(d)
    f = pc.Flow(d, preset)
    with torch.no_grad():
        f.flow.raw.mul_(wmul)
    m = f.flow.ensure_cuda()
    n = 40000
    x = torch.randn(n, d, device="cuda") * scale
    x[:64] = 5.0; x[64:128] = -5.0; x[128:192, ::2] = 5.0; x[192:256] = 0.0
    for inverse in (True, False):
        a, la = torch.empty_like(x), torch.empty(n, device="cuda")
        b, lb = torch.empty_like(x), torch.empty(n, device="cuda")
        m.sweep_tri_into(x, a, la, inverse=inverse)
        old = config.inverse_path
        config.inverse_path = "sweep"
        try:
            m.sweep_into(x, b, lb, inverse=inverse)
        finally:
            config.inverse_path = old
        torch.cuda.synchronize()
        dx = (a - b).abs()
        dl = (la - lb).abs()
        rel = dx / (1 + b.abs())
        print(json.dumps(dict(d=d, flow=preset, input_scale=scale, weight_scale=wmul, inverse=inverse, frac_outside=float((x.abs() > 5).float().mean()),
                              max_rel_dx=float(rel.max()), max_dladj=float(dl.max()), rows_dx_gt_1e3=int((rel.max(1).values > 1e-3).sum()),
                              rows_dladj_gt_1e2=int((dl > 1e-2).sum()), nan_tri=int(torch.isnan(a).any(1).sum()), nan_ffma=int(torch.isnan(b).any(1).sum()))), flush=True)
