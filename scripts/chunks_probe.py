"""Source-chunk count of the fast tiled kernel (LUDVM_FAST_CHUNKS, default 8) at the single-GPU shape (2^20 rows) and at
the per-rank shape of an 8-GPU run (131072 rows of 2^20): ms per launch."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ludvm_b200 import _lib, ops
n = 1 << 20
rng = np.random.default_rng(20260101)
g, x, z = (torch.tensor(a, device="cuda") for a in (rng.standard_normal(n) * 1e-2, rng.uniform(-20, 0, n), rng.uniform(-4, 4, n)))
ctx = _lib.Context(0, torch.cuda.current_stream().cuda_stream)
xo, zo = torch.empty_like(x), torch.empty_like(z)
for chunks in (8, 16, 32):
    os.environ["LUDVM_FAST_CHUNKS"] = str(chunks)
    for rows in (n // 8, n // 4, n):
        best = 1e9
        for _ in range(3 if rows < n else 2):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); ops.selfconv_step(ctx, "fast", g, x, z, 0.065 ** 4, 0.05, xo, zo, row0=0, nrows=rows); e1.record(); torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        print("chunks %2d rows %7d: %8.2f ms  %.4g pairs/s" % (chunks, rows, best, rows * float(n) / best * 1e3), flush=True)
