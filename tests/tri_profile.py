"""A few launches of the tcgen05 block-triangular inverse for ncu (run on the GPU box; not a pytest):
python tests/tri_profile.py D N [flow]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pocomc_b200.flow import Flow
d = int(sys.argv[1]) if len(sys.argv) > 1 else 32
n = int(sys.argv[2]) if len(sys.argv) > 2 else 10000
torch.manual_seed(0)
f = Flow(d, sys.argv[3] if len(sys.argv) > 3 else "maf6")
x = torch.randn(n, d, device="cuda")
z = torch.empty_like(x); l = torch.empty(n, device="cuda")
for _ in range(3):
    f.flow.sweep_tri_into(x, z, l, True)
torch.cuda.synchronize()
