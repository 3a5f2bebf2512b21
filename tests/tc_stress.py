"""Repeatability stress of the tcgen05 forward (not a pytest): the same input must give bit-identical output on
every launch; prints how many launches differ from the first, the worst deviation from the oracle, and how far the
host's plain fp32 CPU GEMMs put the oracle from its fp64-contraction (checker) mode."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
import numpy as np, torch
import flow_ref as F
from pocomc_b200.flow import Flow

d, preset = 32, "maf6"
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 200
for n in (1000, 20000):
    torch.manual_seed(d * 7 + n)
    ref = F.make_flow(d, preset)
    f = Flow(d, preset)
    flat = np.concatenate([p.detach().numpy().reshape(-1) for p in ref.parameters()])
    with torch.no_grad():
        f.flow.raw.copy_(torch.from_numpy(flat).to(f.flow.raw.device))
    x = torch.randn(n, d)
    import zuko.nn as oracle_nn
    with torch.no_grad():
        oracle_nn.MATMUL_FP64 = False
        z_fp32, _ = ref().transform.call_and_ladj(x)          # plain fp32 CPU GEMMs: host dependent (see oracle/zuko/nn.py)
        oracle_nn.MATMUL_FP64 = True
        z_ref, l_ref = ref().transform.call_and_ladj(x)       # checker mode: fp64 contractions
    print(f"n={n} oracle: max |z(fp32 GEMM) - z(fp64 GEMM)| on this host = {float((z_fp32 - z_ref).abs().max()):.3e}", flush=True)
    xd = x.cuda()
    from pocomc_b200 import config
    zs, ls = torch.empty_like(xd), torch.empty(n, device="cuda")
    f.flow.sweep_into(xd, zs, ls, False)
    img = f.flow.packed_tc()
    print(f"n={n} checksums: raw {float(f.flow.raw.double().sum()):.12e} tc image {float(img.double().sum()):.12e} "
          f"|image| {float(img.double().abs().sum()):.12e} z_ref {float(z_ref.double().sum()):.12e} x {float(x.double().sum()):.12e}")
    print(f"n={n} FFMA sweep forward: max |z - oracle| = {float((zs.cpu() - z_ref).abs().max()):.3e}", flush=True)
    z = torch.empty_like(xd); l = torch.empty(n, device="cuda")
    first, bad, worst = None, 0, 0.0
    for r in range(reps):
        z.zero_(); l.zero_()
        f.flow.forward_tc_into(xd, z, l, 3)
        zc = z.cpu()
        err = float((zc - z_ref).abs().max())
        worst = max(worst, err)
        if first is None:
            first = zc.clone(); print(f"n={n} first launch: max |z - oracle| = {err:.3e}", flush=True)
        elif not torch.equal(zc, first):
            bad += 1
            rows = torch.nonzero((zc != first).any(dim=1)).flatten()
            if bad <= 5:
                print(f"  launch {r}: {len(rows)} rows differ (tiles {sorted(set((rows // 128).tolist()))[:10]}), max dev {float((zc - first).abs().max()):.3e}, err vs oracle {err:.3e}", flush=True)
    print(f"n={n}: {bad} of {reps - 1} repeat launches differ from the first; worst error vs oracle {worst:.3e}", flush=True)
