"""Driver for `ncu -k regex:k_exact_rows`: one exact-mode all-pairs evaluation of N vortices (default 2^17)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ludvm_b200 import _lib, ops
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 17
rng = np.random.default_rng(20260101)
g, x, z = (torch.tensor(a, device="cuda") for a in (rng.standard_normal(n) * 1e-2, rng.uniform(-20, 0, n), rng.uniform(-4, 4, n)))
ctx = _lib.Context(0, torch.cuda.current_stream().cuda_stream)
xo, zo, uo, wo = (torch.empty_like(x) for _ in range(4))
for _ in range(2):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); ops.selfconv_step(ctx, "exact", g, x, z, 0.065 ** 4, 0.05, xo, zo, u_out=uo, w_out=wo); e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1)
print("exact N=%d: %.2f ms, %.4g pairs/s" % (n, ms, n * float(n) / ms * 1e3))
