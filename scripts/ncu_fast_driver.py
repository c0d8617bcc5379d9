"""Driver for ncu captures of the fast all-pairs kernels: `nrows` target rows (default 131072 = the per-rank launch of the
8-GPU run) against N = 2^20 sources, two launches.  LUDVM_NO_FUSED=1 selects the partial-sum path."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ludvm_b200 import _lib, ops
nrows = int(sys.argv[1]) if len(sys.argv) > 1 else 131072
mode = sys.argv[2] if len(sys.argv) > 2 else "fast"
n = 1 << 20
rng = np.random.default_rng(20260101)
x, z = (torch.tensor(rng.uniform(a, b, n), device="cuda") for a, b in ((-20, 0), (-4, 4)))
g = torch.tensor(rng.standard_normal(n) * 1e-2, device="cuda")
ctx = _lib.Context(0, torch.cuda.current_stream().cuda_stream)
xo, zo = torch.empty_like(x), torch.empty_like(z)
for _ in range(2):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); ops.selfconv_step(ctx, mode, g, x, z, 0.065 ** 4, 0.05, xo, zo, row0=0, nrows=nrows); e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1)
print("%s rows=%d N=%d: %.2f ms, %.4g pairs/s, plan %s" % (mode, nrows, n, ms, nrows * float(n) / ms * 1e3, ctx.last_plan()))
