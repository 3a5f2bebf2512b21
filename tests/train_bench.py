"""Per-optimiser-step time of Flow.fit's paths (run on the GPU box; not a pytest):
fused kernels in a CUDA graph vs autograd in a CUDA graph vs eager autograd."""
import sys, os, json, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pocomc_b200 import config
from pocomc_b200.flow import Flow, _FitEngine, epoch_batches

profile_only = len(sys.argv) > 1 and sys.argv[1] == "profile"
for d, preset in ((10, "maf6"), (32, "maf6"), (50, "maf6")):
    torch.manual_seed(0)
    x = torch.randn(8192, d, device="cuda")
    w = torch.rand(8192, device="cuda") + 0.1
    for kernels in ("fused", "autograd"):
        if profile_only and kernels != "fused":
            continue
        config.fit_kernels = kernels
        f = Flow(d, preset)
        eng = _FitEngine(f.flow)
        eng.load(x, w)
        eng.reset_optimizer(1e-3, 0.0, 1.0)
        if profile_only:
            eng.loss_and_grad(torch.arange(512), True)
            torch.cuda.synchronize()
            continue
        batches = [torch.arange(i, i + 512) for i in range(0, 8192, 512)]
        eng.run_epoch(batches, 512, True)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        reps = 5
        for _ in range(reps):
            eng.run_epoch(batches, 512, True)
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / (reps * len(batches))
        print(json.dumps(dict(d=d, flow=preset, path=kernels + "+graph", us_per_step=dt * 1e6, loss=float(eng.acc.item()) / 8192)))
    if profile_only:
        continue
    config.fit_path = "eager"
    f = Flow(d, preset)
    t0 = time.perf_counter()
    f.fit(x, weights=w, epochs=3, batch_size=512, shuffle=False, annealing=False)
    torch.cuda.synchronize()
    print(json.dumps(dict(d=d, flow=preset, path="eager autograd (torch AdamW)", us_per_step=(time.perf_counter() - t0) / 48 * 1e6)))
    config.fit_path = "graph"
