"""One treecode evaluation (N = 2^k, order from argv) for `ncu` launch lists / captures."""
import os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ludvm_b200 import _lib, ops
lg = int(sys.argv[1]) if len(sys.argv) > 1 else 20
order = int(sys.argv[2]) if len(sys.argv) > 2 else 18
n = 1 << lg
dev = torch.device("cuda", 0)
ctx = _lib.Context(0, torch.cuda.current_stream(dev).cuda_stream)
rng = np.random.default_rng(20260101)
x = torch.from_numpy(rng.uniform(-20, 0, n)).to(dev)
z = torch.from_numpy(rng.uniform(-4, 4, n)).to(dev)
g = torch.from_numpy(rng.standard_normal(n) * 1e-2).to(dev)
u, w = torch.empty_like(x), torch.empty_like(x)
for _ in range(2):
    st = ops.induced_velocity_tree_device(ctx, g, x, z, x, z, 0.065 ** 4, u, w, order=order, return_stats=True)
torch.cuda.synchronize()
print(st)
