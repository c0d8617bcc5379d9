"""One configs[4] flow-field evaluation through the far field, for ncu launch lists."""
import os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ludvm_b200 import _lib, ops
dev = torch.device("cuda", 0)
ctx = _lib.Context(0, torch.cuda.current_stream(dev).cuda_stream)
rng = np.random.default_rng(20260102)
n = 200000
x = torch.from_numpy(rng.uniform(-20, 0, n)).to(dev); z = torch.from_numpy(rng.uniform(-4, 4, n)).to(dev)
g = torch.from_numpy(rng.standard_normal(n) * 1e-2).to(dev)
x1 = torch.from_numpy(np.arange(-20.48, 0, 0.005)).to(dev); z1 = torch.from_numpy(np.arange(-10.24, 10.24, 0.005)).to(dev)
u, w = (torch.empty((x1.numel(), z1.numel()), dtype=torch.float64, device=dev) for _ in range(2))
order = int(sys.argv[1]) if len(sys.argv) > 1 else 18
for _ in range(2):
    st = ops.flowfield_velocity_tree_device(ctx, g, x, z, 0.065 ** 4, x1, z1, 0, x1.numel(), u, w, 1 / 0.005 ** 2, order=order, return_stats=True)
torch.cuda.synchronize()
print(st)
