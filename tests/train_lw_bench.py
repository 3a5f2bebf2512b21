"""Optimiser-step time of Flow.fit per flow shape: layer-wise kernels (csrc/flow_train_lw.cu) vs torch autograd inside the same
CUDA-graph machinery (run on the GPU box; not a pytest).  python tests/train_lw_bench.py [preset d]..."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pocomc_b200 import config
from pocomc_b200.flow import Flow, _FitEngine, epoch_batches

cases = [(sys.argv[i], int(sys.argv[i + 1])) for i in range(1, len(sys.argv) - 1, 2)] or [("nsf6", 10), ("nsf6", 32), ("nsf3", 50), ("maf6", 100), ("maf6", 200), ("maf6", 32)]
for preset, d in cases:
    rec = dict(flow=preset, d=d, batch=512)
    for mode in ("fused", "autograd"):
        config.fit_kernels = mode
        torch.manual_seed(0)
        f = Flow(d, preset)
        eng = _FitEngine(f.flow)
        x = torch.randn(8192, d, device="cuda")
        w = torch.rand(8192, device="cuda") + 0.1
        eng.load(x, w)
        eng.reset_optimizer(1e-3, 0.0, 1.0)
        batches = epoch_batches(8192, 512, True)
        for _ in range(2):
            eng.run_epoch(batches, 512, True)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        reps = 5
        for _ in range(reps):
            acc = eng.run_epoch(batches, 512, True)
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / (reps * len(batches))
        key = ("layerwise" if eng.layerwise else "fused") if mode == "fused" else "autograd"
        rec[key + "_us_per_step"] = dt * 1e6
        rec[key + "_loss"] = float(acc.item()) / 8192
    print(json.dumps(rec), flush=True)
