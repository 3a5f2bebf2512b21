"""Timing of the tcgen05 dense forward vs the FFMA sweep forward (run on the GPU box; not a pytest)."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pocomc_b200.flow import Flow
from pocomc_b200 import config

def timeit(fn, reps=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps

torch.manual_seed(0)
f = Flow(32, "maf6")
lay = f.flow.layout
dense_flop = 2 * (32 * 128 + 2 * 128 * 128 + 128 * 64) * 6          # per particle, padded dense MLP, one pass
for n in (10_000, 18_944, 100_000, 1_000_000):
    x = torch.randn(n, 32, device="cuda")
    z = torch.empty_like(x); l = torch.empty(n, device="cuda")
    t3 = timeit(lambda: f.flow.forward_tc_into(x, z, l, 3))
    t1 = timeit(lambda: f.flow.forward_tc_into(x, z, l, 1))
    ts = timeit(lambda: f.flow.sweep_into(x, z, l, False), reps=5)
    print(json.dumps(dict(n=n, tc3_ms=t3, tc1_ms=t1, sweep_ms=ts, tc3_issued_tflops=3 * dense_flop * n / t3 / 1e9,
                          tc1_issued_tflops=dense_flop * n / t1 / 1e9, tc3_useful_tflops=dense_flop * n / t3 / 1e9)))
