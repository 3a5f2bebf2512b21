"""Three launches of the flow inverse sweep (default variant) for ncu (run on the GPU box; not a pytest)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pocomc_b200.flow import Flow
n = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
torch.manual_seed(0)
f = Flow(32, "maf6")
x = torch.randn(n, 32, device="cuda")
z = torch.empty_like(x); l = torch.empty(n, device="cuda")
for _ in range(3):
    f.flow.sweep_into(x, z, l, inverse=True)
torch.cuda.synchronize()
