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


def require_equal_blocks(rows: int, group, what: str) -> None:
    """The peer-mapped layers index a neighbour's buffer with this rank's own row count: every rank must hold the same
    number of rows (uneven blocks: use HaloExchanger / ShardedLayer / ShardedSchelling, which carry per-rank bounds)."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return
    mine = torch.tensor([int(rows)], dtype=torch.int64)
    backend = dist.get_backend(group)
    if backend == "nccl":
        mine = mine.cuda()
    everyone = [torch.empty_like(mine) for _ in range(dist.get_world_size(group))]
    dist.all_gather(everyone, mine, group=group)
    counts = [int(t.item()) for t in everyone]
    if len(set(counts)) != 1:
        raise ValueError(f"{what} needs the same number of rows on every rank, got {counts}")


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


class ShardedLayer:
    """Row block of a property layer of a row-sharded grid ("grid stencils: GoL, Schelling happiness,
    layer diffusion" -- one exchange step per stencil).

    Rank k holds rows block_bounds(rows_total, world, k) of a [rows_total, cols] layer of any dtype the layer
    kernels support.  `neighborhood_sum` and `diffuse` exchange `radius` halo rows with the ring neighbours
    (HaloExchanger: NCCL on GPUs, gloo on CPU) and launch the same kernels as the single-GPU path with
    halo_top / halo_bot set; `reduce` is the local reduction kernel + one scalar all-reduce.  `kernels` is the
    namespace providing neighborhood_sum / layer_diffuse / layer_reduce (default mesa_b200.ops; the gloo tests
    inject CPU restatements, there is no CPU path in the product)."""

    _ALLREDUCE = {"sum": "SUM", "count_nonzero": "SUM", "min": "MIN", "max": "MAX"}

    def __init__(self, block: torch.Tensor, torus: bool = False, group=None, kernels=None):
        if block.dim() != 2:
            raise ValueError("ShardedLayer expects a [rows, cols] row block")
        if kernels is None:
            from . import ops as kernels
        self.k = kernels
        self.block = block.contiguous()
        self.torus = bool(torus)
        self.group = group
        self._hx: dict[int, HaloExchanger] = {}
        self._spare = None

    def _halos(self, radius: int):
        hx = self._hx.get(radius)
        if hx is None:
            hx = self._hx[radius] = HaloExchanger(self.block.shape[1], self.block.dtype, radius, self.torus,
                                                  self.block.device, self.group)
        return hx.exchange(self.block)

    def neighborhood_sum(self, radius: int = 1, moore: bool = True, include_center: bool = False) -> torch.Tensor:
        """Neighbourhood sums of this rank's rows (int32 / float64 like ops.neighborhood_sum)."""
        if radius < 1:
            raise ValueError("radius must be at least 1")
        top, bot = self._halos(radius)
        if self.torus and top is None:      # unreachable: a torus always has ring neighbours
            raise RuntimeError("torus without ring neighbours")
        return self.k.neighborhood_sum(self.block, radius, moore, include_center, self.torus, halo_top=top, halo_bot=bot)

    def diffuse(self, rate: float) -> torch.Tensor:
        """One diffusion step of the whole sharded layer; the block is replaced by its new rows."""
        top, bot = self._halos(1)
        if self._spare is None:
            self._spare = torch.empty_like(self.block)
        out = self.k.layer_diffuse(self.block, rate, self.torus, out=self._spare,
                                   halo_top=None if top is None else top[0], halo_bot=None if bot is None else bot[0])
        self.block, self._spare = out, self.block
        return self.block

    def reduce(self, op: str = "sum"):
        """Aggregate over ALL ranks' rows (python int / float)."""
        local = self.k.layer_reduce(self.block, op)
        if dist.is_initialized() and dist.get_world_size(self.group) > 1:
            dist.all_reduce(local, op=getattr(dist.ReduceOp, self._ALLREDUCE[op]), group=self.group)
        return local.item()


class PeerHaloLayer:
    """Row block of a sharded uint8/int8 layer whose halos are READ IN PLACE from the neighbours' HBM.

    Both ping-pong generations of the block live in torch symmetric memory (CUDA peer mappings over
    NVLink 5 / NVSwitch, every peer one hop away), so the stencil kernel receives the neighbours' edge
    rows as plain device pointers (`halo_top` / `halo_bot` of mb_gol_step_u8, mb_schelling_happy_i8):
    no pack/unpack, no NCCL send/recv, no staging copy.  The default generation is one stencil
    launch + one device-side barrier; the optional overlapped order is

        edge kernel (rows 0..r-1 and R-r..R-1 of the new generation, reads the peers' old edges)
        put_signal(up), put_signal(down)          "my new edges are written, your old ones are read"
        interior kernel (rows r..R-r-1, needs no remote data)          <- hides the signal latency
        wait_signal(up), wait_signal(down)

    which also orders the write-after-read on the ping-pong buffers (a peer signals only after the
    kernel that read my old edge rows has finished).  CUDA tensors + NCCL process group only; the
    gloo/CPU path and the fallback use HaloExchanger.
    """

    def __init__(self, rows: int, cols: int, dtype: torch.dtype, radius: int, torus: bool, device, group=None):
        import torch.distributed._symmetric_memory as symm_mem

        self.group = group if group is not None else dist.group.WORLD
        self.rank, self.world = dist.get_rank(self.group), dist.get_world_size(self.group)
        self.rows, self.cols, self.radius, self.torus = int(rows), int(cols), int(radius), bool(torus)
        if self.rows < 2 * self.radius + 1:
            raise ValueError("row block thinner than 2*radius+1")
        require_equal_blocks(self.rows, self.group, "PeerHaloLayer")
        self.buf = symm_mem.empty((2, self.rows, self.cols), dtype=dtype, device=device)
        self.hdl = symm_mem.rendezvous(self.buf, self.group)
        self.up = self._peer(self.rank - 1)
        self.down = self._peer(self.rank + 1)
        shape = (2, self.rows, self.cols)
        self.peer_up = self.hdl.get_buffer(self.up, shape, dtype) if self.up is not None else None
        self.peer_down = self.hdl.get_buffer(self.down, shape, dtype) if self.down is not None else None
        self.cur = 0
        self.hdl.barrier()

    def _peer(self, r: int):
        if 0 <= r < self.world:
            return r
        return r % self.world if self.torus else None

    @property
    def state(self) -> torch.Tensor:
        return self.buf[self.cur]

    def halos(self):
        """(rows above, rows below) of the current generation, living in the neighbours' memory."""
        r, R = self.radius, self.rows
        top = self.peer_up[self.cur, R - r:] if self.peer_up is not None else None
        bot = self.peer_down[self.cur, :r] if self.peer_down is not None else None
        return top, bot

    def step(self, kernel, overlap: bool = False) -> None:
        """Advance one generation.  `kernel(src_rows, dst_rows, halo_top, halo_bot)` runs the stencil on a
        contiguous run of rows given the rows above / below it (None = outside the grid).

        overlap=False (default): ONE launch over the whole block with the neighbours' rows as halo
        pointers, then a device-side barrier over the group (measured 112 us/generation for 16384^2
        per GPU at 2 GPUs against 96 us on one GPU).  overlap=True: edges / signals / interior /
        waits as described in the class docstring; correct, but the six extra small launches cost
        more than the barrier they hide (measured 130-138 us), so it is kept as an option only."""
        r, R = self.radius, self.rows
        src, dst = self.buf[self.cur], self.buf[1 - self.cur]
        top, bot = self.halos()
        if not overlap:
            kernel(src, dst, top, bot)
            self.hdl.barrier()
            self.cur ^= 1
            return
        # edges first: they are what the neighbours wait for
        kernel(src[:r], dst[:r], top, src[r:2 * r])
        kernel(src[R - r:], dst[R - r:], src[R - 2 * r:R - r], bot)
        if self.up is not None:
            self.hdl.put_signal(self.up, 0)
        if self.down is not None:
            self.hdl.put_signal(self.down, 1)
        kernel(src[r:R - r], dst[r:R - r], src[:r], src[R - r:])
        if self.down is not None:
            self.hdl.wait_signal(self.down, 0)   # the rank below signalled its "up" neighbour (me) on channel 0
        if self.up is not None:
            self.hdl.wait_signal(self.up, 1)
        self.cur ^= 1


class FusedPeerGol:
    """Row-sharded Game of Life with the halo hand-shake FUSED into the stencil kernel.

    Like PeerHaloLayer the two ping-pong generations of the block live in symmetric memory and the kernel
    reads the neighbours' edge rows in place over NVLink; unlike it there is no barrier and no exchange at
    all: mb_gol_step_u8_fused waits on / publishes per-neighbour generation counters that also live in
    peer-mapped memory (csrc/gol.cu, PeerSync).  One kernel launch per generation and per GPU."""

    def __init__(self, rows: int, cols: int, torus: bool, device, group=None):
        import torch.distributed._symmetric_memory as symm_mem

        from . import _native as N

        self._N = N
        self.group = group if group is not None else dist.group.WORLD
        self.rank, self.world = dist.get_rank(self.group), dist.get_world_size(self.group)
        self.rows, self.cols, self.torus = int(rows), int(cols), bool(torus)
        self.device = torch.device(device)
        require_equal_blocks(self.rows, self.group, "FusedPeerGol")
        self.buf = symm_mem.empty((2, self.rows, self.cols), dtype=torch.uint8, device=self.device)
        self.hdl = symm_mem.rendezvous(self.buf, self.group)
        self.flags = symm_mem.empty((4,), dtype=torch.int32, device=self.device)   # [wait_top, wait_bot, -, -]
        self.flags.zero_()
        self.fhdl = symm_mem.rendezvous(self.flags, self.group)
        self.local = torch.zeros(4, dtype=torch.int32, device=self.device)          # [edge_done x2, error, -]
        up = self.rank - 1 if self.rank > 0 else (self.world - 1 if self.torus else None)
        down = self.rank + 1 if self.rank < self.world - 1 else (0 if self.torus else None)
        shape = (2, self.rows, self.cols)
        self.peer_up = self.hdl.get_buffer(up, shape, torch.uint8) if up is not None else None
        self.peer_down = self.hdl.get_buffer(down, shape, torch.uint8) if down is not None else None
        self.signal_up = self.fhdl.buffer_ptrs[up] + 4 if up is not None else None        # its wait_bot
        self.signal_down = self.fhdl.buffer_ptrs[down] + 0 if down is not None else None  # its wait_top
        self.cur, self.gen = 0, 0
        torch.cuda.current_stream(self.device).synchronize()
        dist.barrier(group=self.group)

    @property
    def state(self) -> torch.Tensor:
        return self.buf[self.cur]

    def load(self, block: torch.Tensor) -> None:
        """Set the current generation (collective: every rank calls it before stepping)."""
        self.buf[self.cur].copy_(block)
        torch.cuda.current_stream(self.device).synchronize()
        dist.barrier(group=self.group)

    def step(self) -> None:
        N = self._N
        self.gen += 1
        src, dst = self.buf[self.cur], self.buf[1 - self.cur]
        top = self.peer_up[self.cur, self.rows - 1] if self.peer_up is not None else None
        bot = self.peer_down[self.cur, 0] if self.peer_down is not None else None
        f, l = self.flags.data_ptr(), self.local.data_ptr()
        with torch.cuda.device(self.device):
            N.check(N.lib().mb_gol_step_u8_fused(
                src.data_ptr(), N.ptr(top), N.ptr(bot), dst.data_ptr(), self.rows, self.cols, 1,
                f, f + 4, self.signal_up, self.signal_down, l, l + 8, self.gen & 0xFFFFFFFF,
                N.stream_ptr(self.device)))
        self.cur ^= 1

    def check(self) -> None:
        """Raise if a wait timed out (a neighbour never published its edge rows)."""
        if int(self.local[2].item()) != 0:
            raise RuntimeError("FusedPeerGol: timed out waiting for a neighbouring rank")


class DeepGhostLayer:
    """Row block of a sharded layer with DEEP GHOST ZONES: `ghost` rows of each ring neighbour on either side, two
    ping-pong planes.  The extended block is stepped as a clamped grid: the rows next to its outer edge go wrong at one
    row per generation, so the owned rows stay exact for `ghost` generations; only then are the ghost rows refreshed
    from the neighbours -- ONE exchange per `ghost` generations instead of one per launch.  A clamped outer edge of the
    grid gets no ghost rows (cells outside the grid stay dead for ever, ghost rows are ordinary cells).

    transport="peer" (default): the planes live in torch symmetric memory and every rank copies the neighbours' edge
    rows out of their HBM over NVLink with ONE small kernel whose hand-shake is a pair of monotone counters per
    neighbour in peer-mapped memory (csrc/ghost.cu: `mb_ghost_exchange`, `mb_ghost_wait_copied`) -- no barrier, no
    collective, nothing on the host, so a run is a static launch sequence that can be captured in a CUDA graph.
    transport="barrier": the first version, the same copies between two device-wide symmetric-memory barriers.
    transport="p2p": plain tensors, the edge rows travel with send/recv (HaloExchanger; gloo on CPU in the tests)."""

    def __init__(self, rows: int, width: int, dtype: torch.dtype, ghost: int, torus: bool, device, group=None,
                 transport: str = "peer"):
        if transport not in ("peer", "barrier", "p2p"):
            raise ValueError("transport must be 'peer', 'barrier' or 'p2p'")
        self.group = group if group is not None else (dist.group.WORLD if dist.is_initialized() else None)
        self.rank = dist.get_rank(self.group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(self.group) if dist.is_initialized() else 1
        self.rows, self.width, self.torus = int(rows), int(width), bool(torus)
        self.ghost = g = max(1, min(int(ghost), self.rows))
        self.device = torch.device(device)
        up = self.rank - 1 if self.rank > 0 else (self.world - 1 if self.torus else None)
        down = self.rank + 1 if self.rank < self.world - 1 else (0 if self.torus else None)
        shape = (2, self.rows + 2 * g, self.width)
        itemsize = torch.empty((), dtype=dtype).element_size()
        if transport == "peer" and ((g * self.width * itemsize) % 16 or (self.rows * self.width * itemsize) % 16):
            transport = "barrier"   # the copy kernel moves 16-byte vectors
        self.transport = transport
        if transport in ("peer", "barrier"):
            import torch.distributed._symmetric_memory as symm_mem

            require_equal_blocks(self.rows, self.group, "DeepGhostLayer")
            self.buf = symm_mem.empty(shape, dtype=dtype, device=self.device)
            self.buf.zero_()
            self.hdl = symm_mem.rendezvous(self.buf, self.group)
            self.peer_up = self.hdl.get_buffer(up, shape, dtype) if up is not None else None
            self.peer_down = self.hdl.get_buffer(down, shape, dtype) if down is not None else None
            if transport == "peer":
                from . import _native as N

                self._N = N
                self.flags = symm_mem.empty((8,), dtype=torch.int32, device=self.device)
                self.flags.zero_()
                self._fhdl = symm_mem.rendezvous(self.flags, self.group)
                self._up_flags = self._fhdl.buffer_ptrs[up] if up is not None else None
                self._down_flags = self._fhdl.buffer_ptrs[down] if down is not None else None
                self._bytes = g * self.width * itemsize
                self._copied_pending = [False, False]   # a neighbour may still be copying out of this plane
            torch.cuda.current_stream(self.device).synchronize()
            self.hdl.barrier()
            torch.cuda.current_stream(self.device).synchronize()
        else:
            self.buf = torch.zeros(shape, dtype=dtype, device=self.device)
            self.hx = HaloExchanger(self.width, dtype, g, self.torus, self.device, group)
        self.has_up, self.has_down = up is not None, down is not None
        # rows of a plane that are stepped
        self._lo = 0 if self.has_up else g
        self._hi = self.rows + 2 * g if self.has_down else g + self.rows
        self.cur = 0
        self.valid = 0          # generations the ghost rows are still good for
        self.exchanges = 0
        self.waits = 0          # device-side "neighbours have copied" waits launched

    @property
    def state(self) -> torch.Tensor:
        """The owned rows of the current generation."""
        return self.buf[self.cur, self.ghost:self.ghost + self.rows]

    def planes(self):
        """(source, destination) row ranges to step: the owned rows plus the ghost zones that exist."""
        return self.buf[self.cur, self._lo:self._hi], self.buf[1 - self.cur, self._lo:self._hi]

    def invalidate(self) -> None:
        self.valid = 0

    def refresh(self) -> None:
        """Bring the ghost rows of the current generation up to date (every rank calls it at the same generation)."""
        g, R, c = self.ghost, self.rows, self.cur
        mine = self.buf[c]
        if self.transport == "peer":
            N = self._N
            with torch.cuda.device(self.device):
                N.check(N.lib().mb_ghost_exchange(
                    mine[:g].data_ptr() if self.has_up else None,
                    self.peer_up[c, R:R + g].data_ptr() if self.has_up else None,        # the last owned rows of the rank above
                    mine[g + R:].data_ptr() if self.has_down else None,
                    self.peer_down[c, g:2 * g].data_ptr() if self.has_down else None,    # the first owned rows of the rank below
                    self._bytes, self.flags.data_ptr(), self._up_flags, self._down_flags, N.stream_ptr(self.device)))
            self._copied_pending[c] = True
        elif self.transport == "barrier":
            self.hdl.barrier()                       # every rank has finished writing the current generation
            if self.has_up:
                mine[:g].copy_(self.peer_up[c, R:R + g])
            if self.has_down:
                mine[g + R:].copy_(self.peer_down[c, g:2 * g])
            self.hdl.barrier()                       # nobody overwrites a plane a neighbour is still copying from
        else:
            top, bot = self.hx.exchange(mine[g:g + R])
            if self.has_up:
                mine[:g].copy_(top)
            if self.has_down:
                mine[g + R:].copy_(bot)
        self.valid = g
        self.exchanges += 1

    def prepare_write(self) -> None:
        """Call before launching a kernel that writes the destination plane: if the neighbours copied their ghost rows
        out of that plane in the latest exchange, wait (on the device) until they have finished."""
        d = 1 - self.cur
        if self.transport == "peer" and self._copied_pending[d]:
            N = self._N
            with torch.cuda.device(self.device):
                N.check(N.lib().mb_ghost_wait_copied(self.flags.data_ptr(), int(self.has_up), int(self.has_down),
                                                     N.stream_ptr(self.device)))
            self._copied_pending[d] = False
            self.waits += 1

    def check(self) -> None:
        """Raise if a device-side wait timed out (a neighbour never showed up).  Synchronises."""
        if self.transport == "peer" and int(self.flags[6].item()) != 0:
            raise RuntimeError("DeepGhostLayer: timed out waiting for a neighbouring rank")

    def advance(self, generations: int) -> None:
        """The destination plane now holds the state `generations` later."""
        if generations > self.valid:
            raise ValueError("stepped further than the ghost rows allow")
        self.cur ^= 1
        self.valid -= int(generations)


class PeerBitLife:
    """Row-sharded Game of Life on the BIT-PACKED layer (csrc/gol_bits.cu) over a DeepGhostLayer: each rank keeps 64
    rows of each ring neighbour and refreshes them from the neighbours' HBM over NVLink once per 64 generations
    (0.8 % redundant rows on a 16384-row block).  No collective and no barrier on the data path: the refresh is one
    kernel with a flag hand-shake (csrc/ghost.cu), so `run(k)` is a static launch sequence (CUDA-graph capturable)."""

    def __init__(self, rows: int, cols: int, torus: bool, device, generations_per_launch: int = 8,
                 rows_per_segment: int = 0, ghost: int | None = None, group=None, transport: str = "peer",
                 overlap: bool = True):
        from . import ops

        if ghost is None:
            ghost = 64   # measured best at 2 and at 8 GPUs (profiles/): deeper zones cost more rows than they save exchanges

        if cols % 32:
            raise ValueError("the bit-packed layer needs cols to be a multiple of 32")
        self._ops = ops
        self.rows, self.cols, self.torus = int(rows), int(cols), bool(torus)
        self.device = torch.device(device)
        self.generations_per_launch = max(1, min(int(generations_per_launch), ops.GOL_BITS_MAX_GENERATIONS))
        self.rows_per_segment = int(rows_per_segment)
        self.layer = DeepGhostLayer(self.rows, self.cols // 32, torch.int32, max(int(ghost), self.generations_per_launch),
                                    torus, self.device, group, transport=transport)
        self.ghost = self.layer.ghost
        self.launches = 0          # stencil steps (a step with an overlapped refresh is three kernels)
        self.kernel_launches = 0
        self.overlap = overlap
        self._side = self._side2 = None

    @property
    def exchanges(self) -> int:
        return self.layer.exchanges

    @property
    def state(self) -> torch.Tensor:
        """The owned rows of the current generation (bit planes)."""
        return self.layer.state

    def load(self, block: torch.Tensor) -> None:
        """Set the current generation from this rank's uint8 row block (collective)."""
        self._ops.gol_pack(block, out=self.layer.state)
        self.layer.invalidate()

    def store(self, out: torch.Tensor | None = None) -> torch.Tensor:
        return self._ops.gol_unpack(self.layer.state.contiguous(), out=out)

    def run(self, generations: int, even_launches: bool = False) -> None:
        """Advance `generations`.  even_launches=True: end on the plane the run started on (see ops.gol_launch_plan)."""
        refresh_pending = False
        for t in self._ops.gol_launch_plan(generations, self.generations_per_launch, self.layer.valid, self.ghost,
                                           even=even_launches):
            if t == 0:
                # the refresh is folded into the launch that follows it (when the block is tall enough to have an interior)
                if self.overlap and self.layer.transport == "peer" and self.rows >= 4 * self.generations_per_launch:
                    refresh_pending = True
                else:
                    self.layer.refresh()
                continue
            if refresh_pending:
                refresh_pending = False
                self._step_with_refresh(t)
                continue
            self.layer.prepare_write()
            src, dst = self.layer.planes()
            # the extended block is a clamped grid: rows beyond the ghosts are dead (and never reach the owned rows)
            self._ops.gol_step_bits(src, dst, t, torus=False, col_torus=self.torus, rows_per_segment=self.rows_per_segment)
            self.layer.advance(t)
            self.launches += 1
            self.kernel_launches += 1

    def _step_with_refresh(self, t: int) -> None:
        """The first launch after a ghost refresh, with the refresh OVERLAPPED: rows further than `t` rows from the edges
        of the owned block do not depend on the ghost rows within `t` generations, so they are stepped at once on the
        calling stream while a side stream runs the exchange (hand-shake + copy over NVLink, csrc/ghost.cu) and then the
        two edge strips (ghost rows + `t` owned rows each); the streams join afterwards.  Same cells from the same inputs
        as the single launch over the extended block (the strips get the rows across the cut as halo rows)."""
        layer, ops, g, R = self.layer, self._ops, self.ghost, self.rows
        main = torch.cuda.current_stream(self.device)
        if self._side is None:
            # high priority: the exchange and the two small edge launches must not queue up behind the CTAs of the
            # interior launch, which fills the GPU (measured at 8 GPUs: the side work then ran after the interior)
            self._side = torch.cuda.Stream(self.device, priority=-1)
            self._side2 = torch.cuda.Stream(self.device, priority=-1)
        side, side2 = self._side, self._side2
        layer.prepare_write()                       # nobody is still copying out of the plane about to be written
        side.wait_stream(main)                      # the previous launch has written the current plane
        src, dst = layer.buf[layer.cur], layer.buf[1 - layer.cur]
        kw = dict(torus=False, col_torus=self.torus, rows_per_segment=self.rows_per_segment)
        with torch.cuda.stream(side):
            layer.refresh()
            side2.wait_stream(side)                 # the two edge strips are independent of each other
            ops.gol_step_bits(src[layer._lo:g + t], dst[layer._lo:g + t], t, halo_bot=src[g + t:g + 2 * t], **kw)
        with torch.cuda.stream(side2):
            ops.gol_step_bits(src[g + R - t:layer._hi], dst[g + R - t:layer._hi], t, halo_top=src[g + R - 2 * t:g + R - t], **kw)
        ops.gol_step_bits(src[g + t:g + R - t], dst[g + t:g + R - t], t, halo_top=src[g:g + t], halo_bot=src[g + R - t:g + R], **kw)
        main.wait_stream(side)
        main.wait_stream(side2)
        layer.advance(t)
        self.launches += 1
        self.kernel_launches += 3

    def check(self) -> None:
        self.layer.check()
