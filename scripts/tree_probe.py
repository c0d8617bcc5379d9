"""Treecode (csrc/tree.cu) against the all-pairs fused kernel on synthetic clouds: time per evaluation, pairs left, error.
Usage: python scripts/tree_probe.py [log2N ...]   (env PROBE_ORDERS=12,16,18  PROBE_LEAVES=0,256,...)"""
import json, os, sys, time
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ludvm_b200 import _lib, ops

dev = torch.device("cuda", 0)
ctx = _lib.Context(0, torch.cuda.current_stream(dev).cuda_stream)
VC4 = 0.065 ** 4
orders = [int(a) for a in os.environ.get("PROBE_ORDERS", "12,16,18").split(",")]
leaves = [int(a) for a in os.environ.get("PROBE_LEAVES", "0").split(",")]
for lg in [int(a) for a in sys.argv[1:]] or [20]:
    n = 1 << lg
    rng = np.random.default_rng(20260101)
    x = torch.from_numpy(rng.uniform(-20, 0, n)).to(dev)
    z = torch.from_numpy(rng.uniform(-4, 4, n)).to(dev)
    g = torch.from_numpy(rng.standard_normal(n) * 1e-2).to(dev)
    u, w, ur, wr = (torch.empty(n, dtype=torch.float64, device=dev) for _ in range(4))
    # reference rows: the all-pairs fast kernel on a sample of targets (itself checked against the oracle by the tests)
    sel = torch.from_numpy(rng.choice(n, 4096, replace=False)).to(dev)
    xs_, zs_ = x[sel].contiguous(), z[sel].contiguous()
    us, ws = torch.empty(4096, dtype=torch.float64, device=dev), torch.empty(4096, dtype=torch.float64, device=dev)
    ops.induced_velocity_device(ctx, "fast", g, x, z, xs_, zs_, VC4, us, ws)
    den = torch.empty(4096, dtype=torch.float64, device=dev)
    for i0 in range(0, 4096, 64):
        dx, dz = xs_[i0:i0 + 64, None] - x[None, :], zs_[i0:i0 + 64, None] - z[None, :]
        r2 = dx * dx + dz * dz
        den[i0:i0 + 64] = (g.abs()[None, :] * r2.sqrt() / (r2 * r2 + VC4).sqrt()).sum(1) / (2 * np.pi)
    t_all = None
    if lg <= 21:
        ops.induced_velocity_device(ctx, "fast", g, x, z, x, z, VC4, ur, wr); torch.cuda.synchronize()
        t = time.perf_counter(); ops.induced_velocity_device(ctx, "fast", g, x, z, x, z, VC4, ur, wr); torch.cuda.synchronize()
        t_all = time.perf_counter() - t
    for order in orders:
        for leaf in leaves:
            st = ops.induced_velocity_tree_device(ctx, g, x, z, x, z, VC4, u, w, order=order, leaf=leaf, return_stats=True)
            torch.cuda.synchronize()
            best = 1e9
            for _ in range(3):
                t = time.perf_counter(); ops.induced_velocity_tree_device(ctx, g, x, z, x, z, VC4, u, w, order=order, leaf=leaf)
                torch.cuda.synchronize(); best = min(best, time.perf_counter() - t)
            err = float(((u[sel] - us) ** 2 + (w[sel] - ws) ** 2).sqrt().div(den).max())
            print(json.dumps(dict(n=n, order=order, leaf=leaf, ms=best * 1e3, all_pairs_ms=t_all and t_all * 1e3,
                                  speedup=t_all and t_all / best, pairs_left=st["pair_evaluations"] / st["all_pairs"],
                                  eval_pairs_per_s=st["pair_evaluations"] / best, equivalent_pairs_per_s=float(n) * n / best,
                                  max_err_over_sum_abs_terms=err, level=st["leaf_level"], build_ms=st["build_ms"], eval_ms=st["eval_ms"], arena_mb=st["arena_bytes"] / 2 ** 20)), flush=True)
