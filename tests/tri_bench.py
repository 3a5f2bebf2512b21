"""Tensor-core block-triangular sweep (csrc/flow_tri.cu) against the fp32-FMA sweep kernel: max difference and time
per launch (run on the GPU box; not a pytest).  N=10000 D=32 FLOW=maf6 by default."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import pocomc_b200 as pc
from pocomc_b200 import config
torch.manual_seed(0)
d = int(os.environ.get("D", 32))
preset = os.environ.get("FLOW", "maf6")
f = pc.Flow(d, preset)
with torch.no_grad():
    f.flow.raw.mul_(1.2)
m = f.flow
IT = int(os.environ.get("ITER", 20))
for n in [int(v) for v in os.environ.get("N", "10000").split(",")]:
    x = torch.randn(n, d, device="cuda")
    ref, lref = torch.empty_like(x), torch.empty(n, device="cuda")
    out, ladj = torch.empty_like(x), torch.empty(n, device="cuda")
    for inverse in (True, False):
        config.inverse_path = "sweep"
        m.sweep_into(x, ref, lref, inverse=inverse)
        rec = dict(n=n, d=d, flow=preset, inverse=inverse)
        for passes in (3, 1):
            out.zero_(); ladj.zero_()
            m.sweep_tri_into(x, out, ladj, inverse=inverse, passes=passes)
            torch.cuda.synchronize()
            rec[f"maxdiff_p{passes}"] = float((out - ref).abs().max())
            rec[f"maxdiff_ladj_p{passes}"] = float((ladj - lref).abs().max())
            for _ in range(min(3, IT)): m.sweep_tri_into(x, out, ladj, inverse=inverse, passes=passes)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(IT): m.sweep_tri_into(x, out, ladj, inverse=inverse, passes=passes)
            e1.record(); torch.cuda.synchronize()
            rec[f"tri_us_p{passes}"] = e0.elapsed_time(e1) / IT * 1e3
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for _ in range(min(3, IT)): m.sweep_into(x, ref, lref, inverse=inverse)
        e0.record()
        for _ in range(IT): m.sweep_into(x, ref, lref, inverse=inverse)
        e1.record(); torch.cuda.synchronize()
        rec["ffma_us"] = e0.elapsed_time(e1) / IT * 1e3
        print(json.dumps(rec), flush=True)
