import sys, os, time, torch, numpy as np
sys.path.insert(0, '/root/repo')
import pocomc_b200 as pc
torch.manual_seed(0)
n, d = int(os.environ.get("N", 10000)), int(os.environ.get("D", 32))
f = pc.Flow(d, os.environ.get("FLOW", "maf6"))
x = torch.randn(n, d, device='cuda')
out = torch.empty_like(x); ladj = torch.empty(n, device='cuda')
for inv in (True, False):
    for _ in range(5): f.flow.sweep_into(x, out, ladj, inverse=inv)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): f.flow.sweep_into(x, out, ladj, inverse=inv)
    e1.record(); torch.cuda.synchronize()
    print(f"LPP={os.environ.get('PMC_SWEEP_LPP')} n={n} d={d} inverse={inv}: {e0.elapsed_time(e1)/20*1e3:.1f} us")
