"""Row-block sharding of [rows, cols] layers across ranks with r-row halo exchange.

One process per GPU (torch.distributed); rank k owns the contiguous row block k of the global
grid (rows = first index `i` of Grid((d0, d1)), so a block is one contiguous slab).  The only
data-path exchange of the grid stencils (Game of Life, Schelling happiness, diffusion) is the
`radius` edge rows with the two neighbouring ranks; the reference has no counterpart (it is
single-process).  Works with the `gloo` backend on CPU tensors (tests) and `nccl`
on CUDA tensors.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def block_bounds(global_rows: int, world: int, rank: int) -> tuple[int, int]:
    """[lo, hi) rows of `rank`: even split, the first `global_rows % world` ranks get one more."""
    base, extra = divmod(int(global_rows), int(world))
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


class HaloExchanger:
    """Exchanges the top/bottom `radius` rows of a row block with the neighbouring ranks.

    torus=True closes the ring (rank 0 <-> rank world-1); otherwise the end ranks get None
    on their outer side (outside the grid = empty / dead)."""

    def __init__(self, cols: int, dtype: torch.dtype, radius: int, torus: bool, device, group=None):
        self.group = group
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.radius, self.torus = int(radius), bool(torus)
        self.up = self._peer(self.rank - 1)
        self.down = self._peer(self.rank + 1)
        self.halo_top = torch.empty((radius, cols), dtype=dtype, device=device) if self.up is not None else None
        self.halo_bot = torch.empty((radius, cols), dtype=dtype, device=device) if self.down is not None else None

    def _peer(self, r: int):
        if 0 <= r < self.world:
            return r
        return r % self.world if self.torus else None

    def _global(self, r: int) -> int:
        return dist.get_global_rank(self.group, r) if self.group is not None else r

    def start(self, block: torch.Tensor):
        """Post the sends/receives for `block` [rows, cols]; returns the request list."""
        r = self.radius
        if block.shape[0] < r:
            raise ValueError("row block thinner than the halo radius")
        if self.world == 1:
            # single rank: a torus wraps onto itself
            if self.torus:
                self.halo_top.copy_(block[block.shape[0] - r:])
                self.halo_bot.copy_(block[:r])
            return []
        ops = []
        if self.up is not None:
            ops.append(dist.P2POp(dist.isend, block[:r].contiguous(), self._global(self.up), self.group))
            ops.append(dist.P2POp(dist.irecv, self.halo_top, self._global(self.up), self.group))
        if self.down is not None:
            ops.append(dist.P2POp(dist.isend, block[block.shape[0] - r:].contiguous(), self._global(self.down), self.group))
            ops.append(dist.P2POp(dist.irecv, self.halo_bot, self._global(self.down), self.group))
        if self.world == 2 and self.torus:
            # both neighbours are the same peer: order the pairs identically on both ranks
            # (rank 0 talks "down" first, rank 1 talks "up" first) so sends match receives
            if self.rank == 1:
                ops = ops[2:] + ops[:2]
        return dist.batch_isend_irecv(ops) if ops else []

    @staticmethod
    def finish(reqs) -> None:
        for q in reqs:
            q.wait()

    def exchange(self, block: torch.Tensor):
        self.finish(self.start(block))
        return self.halo_top, self.halo_bot
