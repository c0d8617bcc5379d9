"""Fast tiled kernel, default source chunking against LUDVM_FAST_CHUNKS=8 at 2^20 vortices: result difference and ms."""
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
res = {}
for tag in ("8", None):
    if tag: os.environ["LUDVM_FAST_CHUNKS"] = tag
    else: os.environ.pop("LUDVM_FAST_CHUNKS", None)
    u, w = torch.empty_like(x), torch.empty_like(x)
    for _ in range(2):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); ops.selfconv_step(ctx, "fast", g, x, z, 0.065 ** 4, 0.05, xo, zo, u_out=u, w_out=w); e1.record(); torch.cuda.synchronize()
    res[tag] = (u, w, e0.elapsed_time(e1))
d = max(float((res["8"][0] - res[None][0]).abs().max() / res["8"][0].abs().max()), float((res["8"][1] - res[None][1]).abs().max() / res["8"][1].abs().max()))
print("chunks=8: %.2f ms, default: %.2f ms, max|diff|/max|u| = %.3e, finite %s" % (res["8"][2], res[None][2], d, bool(torch.isfinite(res[None][0]).all())))
