"""ncu target: one epoch (16 optimiser steps, batch 512) of the fused Flow.fit path at D = 32 / maf6 (not a pytest).
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 200 --csv python tests/train_profile.py [D]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pocomc_b200.flow import Flow, _FitEngine

d = int(sys.argv[1]) if len(sys.argv) > 1 else 32
torch.manual_seed(0)
x = torch.randn(8192, d, device="cuda")
w = torch.rand(8192, device="cuda") + 0.1
f = Flow(d, "maf6")
eng = _FitEngine(f.flow)
eng.load(x, w)
eng.reset_optimizer(1e-3, 0.0, 1.0)
batches = [torch.arange(i, i + 512) for i in range(0, 8192, 512)]
for _ in range(2):
    eng.run_epoch(batches, 512, True)
torch.cuda.synchronize()
torch.cuda.profiler.start()
eng.run_epoch(batches, 512, True)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
