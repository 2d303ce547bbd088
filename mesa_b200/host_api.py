"""Host-buffer entry points: the call a Mesa user makes when the layer lives in host memory.

The reference keeps property layers as numpy arrays (mesa/discrete_space/grid.py:112).  These functions
take pinned host tensors (or numpy arrays), run the device kernels and return the result in host
memory, pipelining host->device copies, compute and device->host copies over row chunks so that the two
PCIe directions run concurrently."""
from __future__ import annotations

import torch

from . import ops


class GolHostStepper:
    """One Game of Life generation of a uint8 [rows, cols] HOST layer per call (torus).

    The layer is cut into `chunks` row blocks.  Chunk c is uploaded on the copy-in stream, computed on the
    compute stream as soon as chunks c-1, c, c+1 are resident (a chunk needs one row of each neighbour; the
    wrap-around rows of the torus are uploaded first), and downloaded on the copy-out stream, so H2D of
    chunk c+2, the stencil of chunk c and D2H of chunk c-1 overlap."""

    def __init__(self, rows: int, cols: int, device="cuda", chunks: int = 16, ramp: bool = False):
        self.rows, self.cols = int(rows), int(cols)
        self.device = torch.device(device)
        self.d_in = torch.empty((self.rows, self.cols), dtype=torch.uint8, device=self.device)
        self.d_out = torch.empty_like(self.d_in)
        self.s_in, self.s_comp, self.s_out = (torch.cuda.Stream(self.device) for _ in range(3))
        self.bounds = self.chunk_bounds(self.rows, chunks, ramp)
        self.chunks = len(self.bounds)

    @staticmethod
    def chunk_bounds(rows: int, chunks: int, ramp: bool = False) -> list[tuple[int, int]]:
        """Row blocks [lo, hi) covering 0..rows: `chunks` equal blocks, and with `ramp` the first and the last block
        are split again into 1/8, 1/8, 1/4, 1/2 (mirrored at the end) -- the pipeline's fill (the first upload, during
        which nothing can be computed or downloaded) and drain (the last download) shrink from 1/chunks to
        1/(8*chunks) of a one-way transfer.  Measured on B200 / PCIe 5 (16384^2, 16 chunks): 6.36 ms with the ramp,
        6.34 ms without -- the step is bound by the two directions sharing the link (42 GB/s each way concurrently
        against 55 GB/s one way), not by fill / drain, so the ramp is off by default."""
        chunks = max(1, min(int(chunks), rows // 4 if rows >= 8 else 1))
        base, extra = divmod(rows, chunks)
        bounds, lo = [], 0
        for c in range(chunks):
            hi = lo + base + (1 if c < extra else 0)
            bounds.append((lo, hi))
            lo = hi
        if not ramp or chunks < 3 or base < 32:
            return bounds

        def split(lo, hi, fractions):
            n, cuts, at = hi - lo, [], lo
            for f in fractions[:-1]:
                nxt = min(hi, at + max(1, round(n * f)))
                cuts.append((at, nxt))
                at = nxt
            cuts.append((at, hi))
            return [c for c in cuts if c[1] > c[0]]

        head = split(*bounds[0], (0.125, 0.125, 0.25, 0.5))
        tail = split(*bounds[-1], (0.5, 0.25, 0.125, 0.125))
        return head + bounds[1:-1] + tail

    def step(self, h_in: torch.Tensor, h_out: torch.Tensor) -> torch.Tensor:
        if h_in.shape != (self.rows, self.cols) or h_in.dtype != torch.uint8 or h_in.is_cuda or h_out.is_cuda:
            raise ValueError("expected uint8 host tensors of the stepper's shape")
        cur = torch.cuda.current_stream(self.device)
        for s in (self.s_in, self.s_comp, self.s_out):
            s.wait_stream(cur)
        R = self.rows
        ev_in = []
        with torch.cuda.stream(self.s_in):
            # torus wrap rows first: chunk 0 needs row R-1, the last chunk needs row 0
            self.d_in[R - 1].copy_(h_in[R - 1], non_blocking=True)
            for lo, hi in self.bounds:
                self.d_in[lo:hi].copy_(h_in[lo:hi], non_blocking=True)
                e = torch.cuda.Event()
                e.record(self.s_in)
                ev_in.append(e)
        for c, (lo, hi) in enumerate(self.bounds):
            with torch.cuda.stream(self.s_comp):
                self.s_comp.wait_event(ev_in[min(c + 1, self.chunks - 1)])   # chunk c+1 (and everything before) is resident
                ops.gol_step(self.d_in[lo:hi], out=self.d_out[lo:hi], torus=False, col_torus=True,
                             halo_top=self.d_in[(lo - 1) % R], halo_bot=self.d_in[hi % R])
                done = torch.cuda.Event()
                done.record(self.s_comp)
            with torch.cuda.stream(self.s_out):
                self.s_out.wait_event(done)
                h_out[lo:hi].copy_(self.d_out[lo:hi], non_blocking=True)
        cur.wait_stream(self.s_out)
        return h_out


def gol_step_host(h_in: torch.Tensor, h_out: torch.Tensor | None = None, stepper: GolHostStepper | None = None):
    """Convenience wrapper: one generation of a host layer; synchronises before returning."""
    if stepper is None:
        stepper = GolHostStepper(h_in.shape[0], h_in.shape[1])
    if h_out is None:
        h_out = torch.empty_like(h_in).pin_memory()
    stepper.step(h_in, h_out)
    torch.cuda.current_stream(stepper.device).synchronize()
    return h_out


class ShardedGolHostStepper(GolHostStepper):
    """GolHostStepper for a ROW-SHARDED host layer: every rank steps its own [rows, cols] block of a
    [rows * world, cols] grid from / to its own pinned host buffers, and the two halo rows of the block come from the
    ring neighbours (rank k-1's last row, rank k+1's first row) -- the real sharded generation, not a private torus.

    The first and the last row of the block are uploaded first into a small double-buffered tensor in symmetric
    memory; one device-side barrier later every rank reads its neighbours' edge rows in place over NVLink as the
    `halo_top` / `halo_bot` of its first / last chunk.  The chunk pipeline is the parent's."""

    def __init__(self, rows: int, cols: int, device="cuda", chunks: int = 16, torus: bool = True, group=None):
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm_mem

        super().__init__(rows, cols, device, chunks)
        self.group = group if group is not None else dist.group.WORLD
        self.rank, self.world = dist.get_rank(self.group), dist.get_world_size(self.group)
        self.torus = bool(torus)
        self.edges = symm_mem.empty((2, 2, self.cols), dtype=torch.uint8, device=self.device)   # [parity][first, last][cols]
        self.hdl = symm_mem.rendezvous(self.edges, self.group)
        up = self.rank - 1 if self.rank > 0 else (self.world - 1 if self.torus else None)
        down = self.rank + 1 if self.rank < self.world - 1 else (0 if self.torus else None)
        shape = (2, 2, self.cols)
        self.peer_up = self.hdl.get_buffer(up, shape, torch.uint8) if up is not None else None
        self.peer_down = self.hdl.get_buffer(down, shape, torch.uint8) if down is not None else None
        self.parity = 0
        torch.cuda.current_stream(self.device).synchronize()
        self.hdl.barrier()
        torch.cuda.current_stream(self.device).synchronize()

    def step(self, h_in: torch.Tensor, h_out: torch.Tensor) -> torch.Tensor:
        if h_in.shape != (self.rows, self.cols) or h_in.dtype != torch.uint8 or h_in.is_cuda or h_out.is_cuda:
            raise ValueError("expected uint8 host tensors of the stepper's shape")
        cur = torch.cuda.current_stream(self.device)
        R, p = self.rows, self.parity
        self.parity ^= 1
        # edge rows first, then the barrier: past it every rank's edge rows of THIS generation are resident.  (The
        # other parity is what the neighbours read in the previous step; it is rewritten two barriers later.)
        self.edges[p, 0].copy_(h_in[0], non_blocking=True)
        self.edges[p, 1].copy_(h_in[R - 1], non_blocking=True)
        self.hdl.barrier()
        for s in (self.s_in, self.s_comp, self.s_out):
            s.wait_stream(cur)
        top = self.peer_up[p, 1] if self.peer_up is not None else None
        bot = self.peer_down[p, 0] if self.peer_down is not None else None
        ev_in = []
        with torch.cuda.stream(self.s_in):
            for lo, hi in self.bounds:
                self.d_in[lo:hi].copy_(h_in[lo:hi], non_blocking=True)
                e = torch.cuda.Event()
                e.record(self.s_in)
                ev_in.append(e)
        for c, (lo, hi) in enumerate(self.bounds):
            with torch.cuda.stream(self.s_comp):
                self.s_comp.wait_event(ev_in[min(c + 1, self.chunks - 1)])
                ops.gol_step(self.d_in[lo:hi], out=self.d_out[lo:hi], torus=False, col_torus=self.torus,
                             halo_top=self.d_in[lo - 1] if lo > 0 else top,
                             halo_bot=self.d_in[hi] if hi < R else bot)
                done = torch.cuda.Event()
                done.record(self.s_comp)
            with torch.cuda.stream(self.s_out):
                self.s_out.wait_event(done)
                h_out[lo:hi].copy_(self.d_out[lo:hi], non_blocking=True)
        cur.wait_stream(self.s_out)
        return h_out
