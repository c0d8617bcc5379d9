"""CPU test of the multi-rank path: world_size-2 gloo, the per-row kernel replaced by the oracle (tests may use it).
Checks the row partition, the all-gather assembly and that the 2-rank result equals the 1-rank one bitwise."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

N, STEPS, VC, DT = 512, 3, 0.065, 0.05


def _cloud():
    rng = np.random.default_rng(5)
    return rng.standard_normal(N) * 1e-2, rng.uniform(-20, 0, N), rng.uniform(-4, 4, N)


def _oracle_kernel(g, x, z, vc4, dt, row0, nrows, x_out, z_out):
    from oracle import ludvm_oracle as o
    gn, xn, zn = g.numpy(), x.numpy(), z.numpy()
    u, w = o.induced_velocity(gn, xn, zn, xn[row0:row0 + nrows], zn[row0:row0 + nrows], vc4 ** 0.25, nthreads=1)
    x_out[row0:row0 + nrows] = torch.from_numpy(xn[row0:row0 + nrows] + dt * u)
    z_out[row0:row0 + nrows] = torch.from_numpy(zn[row0:row0 + nrows] + dt * w)


def _serial():
    from ludvm_b200.sharded import ShardedSelfConvection
    g, x, z = (torch.from_numpy(a.copy()) for a in _cloud())
    s = ShardedSelfConvection(g, x, z, VC, DT, kernel=_oracle_kernel)
    for _ in range(STEPS):
        xs, zs = s.step()
    return xs.numpy().copy(), zs.numpy().copy()


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from ludvm_b200.sharded import ShardedSelfConvection, shard_bounds
    g, x, z = (torch.from_numpy(a.copy()) for a in _cloud())
    s = ShardedSelfConvection(g, x, z, VC, DT, kernel=_oracle_kernel)
    assert (s.row0, s.nrows) == shard_bounds(N, world, rank) == (rank * N // world, N // world)
    for _ in range(STEPS):
        xs, zs = s.step()
    q.put((rank, xs.numpy().copy(), zs.numpy().copy()))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_rows_equal_serial(oracle):
    with socket.socket() as sk:
        sk.bind(("127.0.0.1", 0))
        port = sk.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    xs, zs = _serial()
    for _, x2, z2 in res:
        assert np.array_equal(x2, xs) and np.array_equal(z2, zs)


def test_shard_bounds_rejects_ragged():
    from ludvm_b200.sharded import shard_bounds
    with pytest.raises(ValueError):
        shard_bounds(10, 4, 0)
    assert shard_bounds(16, 4, 3) == (12, 4)
