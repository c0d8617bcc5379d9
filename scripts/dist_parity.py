"""Multi-GPU parity: the G-rank row-sharded self-convection equals the 1-rank result bitwise (run under torchrun)."""
import os, sys
import numpy as np
import torch
import torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ludvm_b200 import _lib, ops
from ludvm_b200.sharded import ShardedSelfConvection

local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
rank, world = dist.get_rank(), dist.get_world_size()
ctx = _lib.Context(local, torch.cuda.current_stream().cuda_stream)
rng = np.random.default_rng(20260101)
ok = True
for n, mode, steps, transport in ((65536, "fast", 3, "p2p"), (65536, "fast", 3, "nccl"), (16384, "exact", 2, "p2p"),
                                  (65536, "fp32", 2, "nccl"), (131072, "tree", 2, "nccl")):
    x_h, z_h, g_h = rng.uniform(-20, 0, n), rng.uniform(-4, 4, n), rng.standard_normal(n) * 1e-2
    g, x, z = (torch.tensor(a, device=dev) for a in (g_h, x_h, z_h))
    sc = ShardedSelfConvection(g, x.clone(), z.clone(), 0.065, 0.05, mode=mode, ctx=ctx,
                               transport=transport if world > 1 else "auto")
    for _ in range(steps):
        xs, zs = sc.step()
    # single-rank evaluation of the same steps on this GPU
    xa, za = x.clone(), z.clone()
    xb, zb = torch.empty_like(x), torch.empty_like(z)
    for _ in range(steps):
        if mode == "tree":
            ops.selfconv_step_tree(ctx, g, xa, za, 0.065 ** 4, 0.05, xb, zb, order=18)
        else:
            ops.selfconv_step(ctx, mode, g, xa, za, 0.065 ** 4, 0.05, xb, zb)
        xa, xb = xb, xa
        za, zb = zb, za
    torch.cuda.synchronize()
    same = bool(torch.equal(xs, xa) and torch.equal(zs, za))
    ok &= same
    print("rank %d/%d n=%d mode=%s steps=%d transport=%s bitwise_equal_to_single_rank=%s"
          % (rank, world, n, mode, steps, sc.transport, same), flush=True)
flag = torch.tensor([1 if ok else 0], device=dev)
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if int(flag.item()) == 1 else 1)
