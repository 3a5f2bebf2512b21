"""Dev helper: time per Flow.inverse / forward launch of spline flows (fp32-FMA sweep with the RQS head) next to the affine
flows of the same shape (tcgen05 block-triangular sweep).  Not a pytest."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import pocomc_b200 as pc
from pocomc_b200 import made_layout as ML
torch.manual_seed(0)
IT = 20
for d, n in ((10, 256), (10, 1000), (32, 10000), (50, 50000), (100, 50000)):
    for preset in ("nsf6", "maf6"):
        f = pc.Flow(d, preset)
        m = f.flow.ensure_cuda()
        lay = m.layout
        x = torch.randn(n, d, device="cuda") * 0.5
        out, ladj = torch.empty_like(x), torch.empty(n, device="cuda")
        rec = dict(flow=preset, d=d, n=n, H=lay.n_hidden, stream=bool(ML.stream_supported(d, lay.n_hidden, lay.n_layers, lay.kind)))
        for inverse in (True, False):
            for _ in range(3): m.sweep_into(x, out, ladj, inverse=inverse)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(IT): m.sweep_into(x, out, ladj, inverse=inverse)
            e1.record(); torch.cuda.synchronize()
            rec["inverse_us" if inverse else "forward_us"] = round(e0.elapsed_time(e1) / IT * 1e3, 1)
        print(json.dumps(rec), flush=True)
