"""Multi-GPU runs of the two BASELINE.json configurations that shard with NO data-path collective:
configs[3] parameter sweep (cases split over ranks) and configs[4] flow-field grid (x-rows split over ranks, one halo
row per side recomputed locally for the vorticity stencil).  Launch under torchrun; rank 0 prints one JSON line per
config with the max-over-ranks time; each rank checks its slice against the single-GPU evaluation of sampled entries.
    python -m torch.distributed.run --nnodes=1 --nproc-per-node G --master-addr 127.0.0.1 scripts/dist_configs.py"""
import json, os, sys, time
import numpy as np
import torch
import torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ludvm_b200 import _lib, sweep
from ludvm_b200.sharded import case_slice, grid_slab

local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
rank, world = dist.get_rank(), dist.get_world_size()
ctx = _lib.Context(local, torch.cuda.current_stream().cuda_stream)
L = _lib.load()
README = dict(t0=0, tf=20, dt=5e-2, chord=1, rho=1.225, Uinf=1, Npoints=81, Ncoeffs=30, LESPcrit=0.2, Naca="0012")
ncases_side = int(os.environ.get("SWEEP_SIDE", "64"))


def maxtime(seconds):
    t = torch.tensor([seconds], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


# ---- configs[3]: sweep, cases split evenly over the ranks
cases = sweep.lespcrit_k_grid(np.linspace(0.1, 0.4, ncases_side), np.linspace(0.1, 1.0, ncases_side), **README)
sl = case_slice(len(cases), world, rank)
sweep.run_sweep(cases[:2], mode="exact", ctx=ctx)                  # warm-up: module load, kernel attributes
for mode in ("exact", "fast"):
    dist.barrier(); torch.cuda.synchronize()
    t = time.perf_counter(); res = sweep.run_sweep(cases, mode=mode, ctx=ctx, case_slice=sl); dt = time.perf_counter() - t
    tmax = maxtime(res["timing"]["ludvm_sweep_run_s"])
    chk = float(np.sum(res["Cl"][:, -1]))                           # checksum of the slice, summed over ranks below
    c = torch.tensor([chk], dtype=torch.float64, device=dev); dist.all_reduce(c)
    if rank == 0:
        print(json.dumps({"config": "configs[3] sweep", "mode": mode, "n_gpus": world, "cases": len(cases), "steps_per_case": 400,
                          "seconds_max_over_ranks": tmax, "case_steps_per_s": len(cases) * 400 / tmax,
                          "collective": "none", "sum_Cl_last": float(c.item())}), flush=True)

# ---- configs[4]: flow-field 4096 x 4096 grid, 200k sources, x-rows split over ranks (+1 halo row per side)
rng = np.random.default_rng(20260102)
nsrc = 200000
xh, zh, gh = rng.uniform(-20, 0, nsrc), rng.uniform(-4, 4, nsrc), rng.standard_normal(nsrc) * 1e-2
x1, z1 = np.arange(-20.48, 0, 0.005), np.arange(-10.24, 10.24, 0.005)
nx, nz = len(x1), len(z1)
r0, r1, h0, h1 = grid_slab(nx, world, rank)                         # owned rows, slab with halo
g, xs, zs, X1, Z1 = (torch.tensor(a, device=dev) for a in (gh, xh, zh, x1, z1))
u = torch.empty((h1 - h0, nz), dtype=torch.float64, device=dev); w = torch.empty_like(u); ome = torch.empty_like(u)
X1s = X1[h0:h1].contiguous()
for rep in range(2):
    dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    _lib.check(L.ludvm_flowfield_velocity(ctx.handle, _lib.FAST_F64, g.data_ptr(), xs.data_ptr(), zs.data_ptr(), nsrc, None, None, None, 0,
                                          0.065 ** 4, X1.data_ptr(), nx, Z1.data_ptr(), nz, h0, h1 - h0, u.data_ptr(), w.data_ptr(), _lib.PTR_DEVICE))
    _lib.check(L.ludvm_flowfield_vorticity(ctx.handle, X1s.data_ptr(), h1 - h0, Z1.data_ptr(), nz, u.data_ptr(), w.data_ptr(), 1, ome.data_ptr(), _lib.PTR_DEVICE))
    e1.record(); e1.synchronize()
    ms = maxtime(e0.elapsed_time(e1))
pairs = float(nx) * nz * nsrc
# parity of the slab against a direct evaluation of 64 sampled points on this GPU
pick = np.random.default_rng(11 + rank).choice((r1 - r0) * nz, 64, replace=False)
ii, jj = r0 + pick // nz, pick % nz
from ludvm_b200 import ops
xp, zp = torch.tensor(x1[ii], device=dev), torch.tensor(z1[jj], device=dev)
ur, wr = torch.empty(64, dtype=torch.float64, device=dev), torch.empty(64, dtype=torch.float64, device=dev)
_lib.check(L.ludvm_induced_velocity(ctx.handle, _lib.FAST_F64, g.data_ptr(), nsrc, xs.data_ptr(), zs.data_ptr(), None, 0.065 ** 4, nsrc,
                                    xp.data_ptr(), zp.data_ptr(), 64, ur.data_ptr(), wr.data_ptr(), _lib.PTR_DEVICE))
torch.cuda.synchronize()
err = float(torch.max(torch.abs(u[torch.tensor(ii - h0, device=dev), torch.tensor(jj, device=dev)] - ur)) / torch.max(torch.abs(ur)))
e = torch.tensor([err], dtype=torch.float64, device=dev); dist.all_reduce(e, op=dist.ReduceOp.MAX)
if rank == 0:
    print(json.dumps({"config": "configs[4] flowfield 4096x4096 x 200k", "mode": "fast", "n_gpus": world, "ms_max_over_ranks": ms,
                      "pairs_per_s": pairs / (ms * 1e-3), "collective": "none (x-row slabs + 1 halo row per side)",
                      "max_rel_err_sampled_points_vs_direct": float(e.item())}), flush=True)
dist.barrier()
dist.destroy_process_group()
