"""Target-row sharded self-convection of one vortex cloud over the ranks of a torch.distributed group
(BASELINE.json configs[2]; the convection phase LUDVM.py:1095-1127 for a free cloud).

Every rank holds the full source state (32 B/vortex), evaluates the rows [rank*N/G, (rank+1)*N/G) against all N
sources with the CUDA kernel behind `ludvm_selfconv_step`, and the updated positions are all-gathered (16
B/vortex/step).  Row sums do not depend on the shard, so the G-rank result is bitwise equal to the 1-rank one.
The per-row kernel is injectable so the sharding/assembly logic can be exercised on CPU tensors with gloo."""
import torch
import torch.distributed as dist

from . import ops


def shard_bounds(n, world, rank):
    """Contiguous, equal row shards (the all-gather needs equal counts)."""
    if n % world:
        raise ValueError("number of vortices (%d) must be a multiple of the number of ranks (%d)" % (n, world))
    per = n // world
    return rank * per, per


class ShardedSelfConvection:
    def __init__(self, g, x, z, v_core, dt, mode="fast", ctx=None, group=None, kernel=None):
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.g, self.x, self.z = g, x, z
        self.n = x.numel()
        self.vc4, self.dt, self.mode, self.ctx = float(v_core) ** 4, float(dt), mode, ctx
        self.row0, self.nrows = shard_bounds(self.n, self.world, self.rank)
        self._xs, self._zs = torch.empty_like(x), torch.empty_like(z)   # shard rows land here
        self._xn, self._zn = torch.empty_like(x), torch.empty_like(z)   # gathered next state
        self._kernel = kernel or self._cuda_kernel

    def _cuda_kernel(self, g, x, z, vc4, dt, row0, nrows, x_out, z_out):
        ops.selfconv_step(self.ctx, self.mode, g, x, z, vc4, dt, x_out, z_out, row0=row0, nrows=nrows)

    def step(self):
        """One forward-Euler step of the whole cloud; returns the new (x, z) (full length on every rank)."""
        if self.world == 1:
            self._kernel(self.g, self.x, self.z, self.vc4, self.dt, 0, self.n, self._xn, self._zn)
        else:
            r0, nr = self.row0, self.nrows
            self._kernel(self.g, self.x, self.z, self.vc4, self.dt, r0, nr, self._xs, self._zs)
            dist.all_gather_into_tensor(self._xn, self._xs[r0:r0 + nr], group=self.group)
            dist.all_gather_into_tensor(self._zn, self._zs[r0:r0 + nr], group=self.group)
        self.x, self._xn = self._xn, self.x
        self.z, self._zn = self._zn, self.z
        return self.x, self.z
