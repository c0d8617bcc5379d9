"""First-contact probe for the GPU box: FP64/FP32 FMA rates and all-pairs throughput at a few sizes."""
import json
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from ludvm_b200 import _lib, ops

ctx = _lib.Context(0, torch.cuda.current_stream().cuda_stream)
out = {"gpu": torch.cuda.get_device_name(0)}
out["dfma_per_s"] = ctx.fp64_fma_rate(300.0)
out["ffma_per_s"] = ctx.fp32_fma_rate(300.0)
print(json.dumps(out), flush=True)
rng = np.random.default_rng(20260101)
for n, modes in [(16384, ("exact", "fast", "fp32")), (131072, ("exact", "fast", "fp32")), (1 << 20, ("fast", "fp32"))]:
    x = torch.tensor(rng.uniform(-20, 0, n), device="cuda")
    z = torch.tensor(rng.uniform(-4, 4, n), device="cuda")
    g = torch.tensor(rng.standard_normal(n) * 1e-2, device="cuda")
    xo, zo = torch.empty_like(x), torch.empty_like(z)
    for mode in modes:
        for rep in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            ops.selfconv_step(ctx, mode, g, x, z, 0.065 ** 4, 0.05, xo, zo)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1)
        pps = n * n / (ms * 1e-3)
        print(json.dumps({"n": n, "mode": mode, "ms": ms, "pairs_per_s": pps,
                          "frac_dfma_slots": pps * 13 / out["dfma_per_s"]}), flush=True)
