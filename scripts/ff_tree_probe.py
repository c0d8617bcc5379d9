"""configs[4] through the hierarchical far field: leaf population / order sweep (time, pairs left, error at sampled points)."""
import json, os, sys, time
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
nx, nz = x1.numel(), z1.numel()
u, w = (torch.empty((nx, nz), dtype=torch.float64, device=dev) for _ in range(2))
vc4 = 0.065 ** 4
ii, jj = torch.from_numpy(rng.integers(0, nx, 1024)).to(dev), torch.from_numpy(rng.integers(0, nz, 1024)).to(dev)
xs_, zs_ = x1[ii].contiguous(), z1[jj].contiguous()
us, ws = (torch.empty(1024, dtype=torch.float64, device=dev) for _ in range(2))
ops.induced_velocity_device(ctx, "fast", g, x, z, xs_, zs_, vc4, us, ws)
dx, dz = xs_[:, None] - x[None, :], zs_[:, None] - z[None, :]
r2 = dx * dx + dz * dz
den = (g.abs()[None, :] * r2.sqrt() / (r2 * r2 + vc4).sqrt()).sum(1) / (2 * np.pi)
for order in (18, 14):
    for leaf in (0, 16, 32, 64, 128, 256, 722):
        st = ops.flowfield_velocity_tree_device(ctx, g, x, z, vc4, x1, z1, 0, nx, u, w, 1 / 0.005 ** 2, order=order, leaf=leaf, return_stats=True)
        torch.cuda.synchronize()
        best = 1e9
        for _ in range(2):
            t = time.perf_counter(); ops.flowfield_velocity_tree_device(ctx, g, x, z, vc4, x1, z1, 0, nx, u, w, 1 / 0.005 ** 2, order=order, leaf=leaf)
            torch.cuda.synchronize(); best = min(best, time.perf_counter() - t)
        err = float(((u[ii, jj] - us) ** 2 + (w[ii, jj] - ws) ** 2).sqrt().div(den).max())
        print(json.dumps(dict(order=order, leaf=leaf, ms=best * 1e3, level=st["leaf_level"], pairs_left=st["pair_evaluations"] / st["all_pairs"],
                              build_ms=st["build_ms"], eval_ms=st["eval_ms"], err=err)), flush=True)
