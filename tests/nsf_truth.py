"""Dev helper: which spline sweep is closer to the truth where the two kernels disagree?  Truth = the oracle flow evaluated in
float64 (same fp32 weights) on the CPU; compared: the tcgen05 block-triangular sweep (lean spline head; run again with
PMC_B200_LIBPATH pointing at a -DPMC_TRI_RQS_REFERENCE_HEAD build for the reference-ordered head) and the fp32-FMA sweep."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle")]
import numpy as np, torch
import flow_ref as F
import pocomc_b200 as pc
from pocomc_b200 import config

for d, preset, scale, wmul in ((10, "nsf6", 1.0, 2.0), (12, "nsf3", 6.0, 1.5), (10, "nsf6", 1.5, 1.0), (32, "nsf6", 1.0, 1.6)):
    torch.manual_seed(d)
    f = pc.Flow(d, preset)
    with torch.no_grad():
        f.flow.raw.mul_(wmul)
    m = f.flow.ensure_cuda()
    ref = F.make_flow(d, preset).double()
    params = []
    for t in range(m.layout.n_transforms):
        for w, b in m.transform_params(t):
            params += [w.detach().cpu().numpy(), b.detach().cpu().numpy()]
    F.load_params(ref, params)
    n = 6000 if d <= 12 else 1500
    x = torch.randn(n, d) * scale
    with torch.no_grad():
        z64, l64 = ref().transform.call_and_ladj(x.double())
        z32 = z64.float()
        xi64, li64 = ref().transform.inv.call_and_ladj(z32.double())
    xd, zd = x.cuda(), z32.cuda()
    for inverse, src, truth, ltruth in ((False, xd, z64, l64), (True, zd, xi64, li64)):
        rec = dict(d=d, flow=preset, input_scale=scale, weight_scale=wmul, inverse=inverse, lib=os.path.basename(os.environ.get("PMC_B200_LIBPATH", "default")))
        for name in ("tri", "ffma"):
            out, la = torch.empty_like(src), torch.empty(n, device="cuda")
            if name == "tri":
                m.sweep_tri_into(src, out, la, inverse=inverse)
            else:
                old, config.inverse_path = config.inverse_path, "sweep"
                try:
                    m.sweep_into(src, out, la, inverse=inverse)
                finally:
                    config.inverse_path = old
            e = ((out.cpu().double() - truth).abs() / (1 + truth.abs())).max(1).values.numpy()
            el = (la.cpu().double() - ltruth).abs().numpy()
            rec[name] = dict(err_x_p50=float(np.median(e)), err_x_p999=float(np.quantile(e, 0.999)), err_x_max=float(e.max()),
                             err_ladj_p50=float(np.median(el)), err_ladj_p999=float(np.quantile(el, 0.999)), err_ladj_max=float(el.max()),
                             rows_x_gt_5e4=int((e > 5e-4).sum()), rows_ladj_gt_5e3=int((el > 5e-3).sum()))
        print(json.dumps(rec), flush=True)
