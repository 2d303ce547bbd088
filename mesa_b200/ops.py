"""Thin tensor-level wrappers over the C ABI (include/mesa_b200.h).

Every function takes CUDA tensors, enqueues on torch's current stream and returns without
synchronising.  There is no CPU path: CPU tensors raise.
"""
from __future__ import annotations

import torch

from . import _native as N


def _need_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise RuntimeError("mesa_b200 runs on CUDA tensors only (no CPU fallback)")


def _c(t: torch.Tensor) -> torch.Tensor:
    if not t.is_contiguous():
        raise ValueError("expected a contiguous tensor")
    return t


def gol_step(state: torch.Tensor, out: torch.Tensor | None = None, torus: bool = True,
             halo_top: torch.Tensor | None = None, halo_bot: torch.Tensor | None = None,
             col_torus: bool | None = None) -> torch.Tensor:
    """One Game of Life generation of a uint8 [rows, cols] layer (agents.py:33-55 for all cells).

    `torus=True` wraps both directions like OrthogonalMooreGrid(torus=True).  Explicit
    `halo_top/halo_bot` rows (row-sharded grids) override the row wrap.
    """
    _need_cuda(state, out, halo_top, halo_bot)
    if state.dtype != torch.uint8 or state.dim() != 2:
        raise ValueError("gol_step expects a uint8 [rows, cols] tensor")
    _c(state)
    rows, cols = state.shape
    if out is None:
        out = torch.empty_like(state)
    if col_torus is None:
        col_torus = torus
    if halo_top is None and halo_bot is None and torus and rows > 0:
        if rows < 3:
            raise ValueError("torus needs rows >= 3")
        halo_top, halo_bot = state[rows - 1], state[0]
    with torch.cuda.device(state.device):
        N.check(N.lib().mb_gol_step_u8(N.ptr(state), N.ptr(halo_top), N.ptr(halo_bot), N.ptr(_c(out)),
                                       rows, cols, int(bool(col_torus)), N.stream_ptr(state.device)))
    return out


GOL_BITS_MAX_GENERATIONS = 8


def gol_pack(state: torch.Tensor, out: torch.Tensor | None = None) -> torch.Tensor:
    """uint8 0/1 [rows, cols] layer -> int32 [rows, cols/32] bit planes (bit k of word w = column 32w+k)."""
    _need_cuda(state, out)
    if state.dtype != torch.uint8 or state.dim() != 2:
        raise ValueError("gol_pack expects a uint8 [rows, cols] tensor")
    rows, cols = state.shape
    if cols % 32:
        raise ValueError("the bit-packed layer needs cols to be a multiple of 32")
    if out is None:
        out = torch.empty((rows, cols // 32), dtype=torch.int32, device=state.device)
    with torch.cuda.device(state.device):
        N.check(N.lib().mb_gol_pack_u8(N.ptr(_c(state)), N.ptr(_c(out)), rows, cols, N.stream_ptr(state.device)))
    return out


def gol_unpack(bits: torch.Tensor, out: torch.Tensor | None = None) -> torch.Tensor:
    """int32 [rows, cols/32] bit planes -> uint8 0/1 [rows, cols] layer."""
    _need_cuda(bits, out)
    if bits.dtype != torch.int32 or bits.dim() != 2:
        raise ValueError("gol_unpack expects an int32 [rows, cols/32] tensor")
    rows, wpr = bits.shape
    if out is None:
        out = torch.empty((rows, wpr * 32), dtype=torch.uint8, device=bits.device)
    if out.dtype != torch.uint8 or tuple(out.shape) != (rows, wpr * 32):
        raise ValueError("gol_unpack: out must be uint8 [rows, 32 * words]")
    with torch.cuda.device(bits.device):
        N.check(N.lib().mb_gol_unpack_u8(N.ptr(_c(bits)), N.ptr(_c(out)), rows, wpr * 32, N.stream_ptr(bits.device)))
    return out


def gol_step_bits(bits: torch.Tensor, out: torch.Tensor, generations: int = 1, torus: bool = True,
                  halo_top: torch.Tensor | None = None, halo_bot: torch.Tensor | None = None,
                  col_torus: bool | None = None, rows_per_segment: int = 0) -> torch.Tensor:
    """`generations` (1..8) Game of Life generations of a bit-packed layer in ONE launch (csrc/gol_bits.cu).

    `halo_top/halo_bot` are [generations, words] packed rows of the neighbouring row blocks; a single-GPU torus
    takes them from the opposite edge of `bits` itself."""
    _need_cuda(bits, out, halo_top, halo_bot)
    if bits.dtype != torch.int32 or bits.dim() != 2 or out.dtype != torch.int32 or out.shape != bits.shape:
        raise ValueError("gol_step_bits expects int32 [rows, cols/32] tensors")
    rows, wpr = bits.shape
    generations = int(generations)
    if col_torus is None:
        col_torus = torus
    if halo_top is None and halo_bot is None and torus and rows > 0:
        if rows < 3 or rows < generations:
            raise ValueError("torus needs rows >= max(3, generations)")
        halo_top, halo_bot = bits[rows - generations:], bits[:generations]
    for h in (halo_top, halo_bot):
        if h is not None and (h.dtype != torch.int32 or tuple(h.shape) != (generations, wpr)):
            raise ValueError("halo rows must be int32 [generations, cols/32]")
    with torch.cuda.device(bits.device):
        N.check(N.lib().mb_gol_step_bits(N.ptr(_c(bits)), N.ptr(halo_top), N.ptr(halo_bot), N.ptr(_c(out)), rows,
                                         wpr * 32, int(bool(col_torus)), generations, int(rows_per_segment),
                                         N.stream_ptr(bits.device)))
    return out


def gol_launch_plan(generations: int, per_launch: int, valid: int | None = None, ghost: int | None = None,
                    even: bool = False) -> list[int]:
    """Generations of each launch of a `generations`-long run: at most `per_launch` per launch and, on a layer with deep
    ghost zones, never past the `valid` generations the ghost rows are good for (0 = refresh first, then `ghost` more).
    A zero entry means "refresh the ghost rows here".  even=True splits the longest launch when the number of launches
    is odd, so that the run ends on the ping-pong plane it started on (a CUDA graph of the run can then be replayed)."""
    plan, left = [], int(generations)
    while left > 0:
        if valid is not None and valid == 0:
            plan.append(0)
            valid = ghost
        t = min(left, per_launch) if valid is None else min(left, per_launch, valid)
        plan.append(t)
        left -= t
        if valid is not None:
            valid -= t
    if even and sum(1 for t in plan if t) % 2:
        # split the LONGEST launch (the last one among equals): two launches of 4 generations cost less than a pair of
        # 2-generation ones (the per-generation cost falls with the number of generations chained in a launch)
        best = max(plan)
        if best < 2:
            raise ValueError("a run of single-generation launches cannot be made even")
        i = len(plan) - 1 - plan[::-1].index(best)
        plan[i:i + 1] = [best - best // 2, best // 2]
    return plan


class BitLife:
    """Game of Life state held bit-packed on one GPU (two ping-pong planes), stepped several generations per
    launch.  `load` / `store` convert from / to the uint8 property layer."""

    def __init__(self, rows: int, cols: int, device, torus: bool = True, generations_per_launch: int = 8,
                 rows_per_segment: int = 0):
        if cols % 32:
            raise ValueError("the bit-packed layer needs cols to be a multiple of 32")
        if torus and (rows < 3 or cols < 3):
            raise ValueError("torus needs rows, cols >= 3")
        self.rows, self.cols, self.torus = int(rows), int(cols), bool(torus)
        self.device = torch.device(device)
        self.rows_per_segment = int(rows_per_segment)
        t = max(1, min(int(generations_per_launch), GOL_BITS_MAX_GENERATIONS))
        self.generations_per_launch = min(t, self.rows) if torus else t
        self.cur = torch.zeros((self.rows, self.cols // 32), dtype=torch.int32, device=self.device)
        self.nxt = torch.empty_like(self.cur)
        self.launches = 0

    def load(self, state: torch.Tensor) -> None:
        gol_pack(state, out=self.cur)

    def store(self, out: torch.Tensor | None = None) -> torch.Tensor:
        return gol_unpack(self.cur, out=out)

    def run(self, generations: int, even_launches: bool = False) -> None:
        for t in gol_launch_plan(generations, self.generations_per_launch, even=even_launches):
            gol_step_bits(self.cur, self.nxt, t, torus=self.torus, rows_per_segment=self.rows_per_segment)
            self.cur, self.nxt = self.nxt, self.cur
            self.launches += 1


def happy_threshold_table(homophily: float, radius: int):
    """min_similar[valid] evaluated with Python floats exactly like agents.py:31-39:
    `frac = similar / valid if valid > 0 else 0.0; happy = not (frac < homophily)`."""
    import numpy as np

    nmax = (2 * radius + 1) ** 2 - 1
    tbl = np.empty(nmax + 1, dtype=np.uint8)
    for valid in range(nmax + 1):
        t = valid + 1  # never happy
        for similar in range(valid + 1):
            frac = similar / valid if valid > 0 else 0.0
            if not (frac < homophily):
                t = similar
                break
        tbl[valid] = min(t, 255)
    return tbl


def schelling_happy(type_layer: torch.Tensor, homophily: float, radius: int = 1, torus: bool = False,
                    out: torch.Tensor | None = None, count: torch.Tensor | None = None,
                    halo_top: torch.Tensor | None = None, halo_bot: torch.Tensor | None = None,
                    binary_types: bool = True, table=None) -> tuple[torch.Tensor, torch.Tensor]:
    """assign_state for every agent of an int8 type layer (-1 empty).  Returns (happy layer, count[1])."""
    _need_cuda(type_layer, out, count, halo_top, halo_bot)
    if type_layer.dtype != torch.int8 or type_layer.dim() != 2:
        raise ValueError("schelling_happy expects an int8 [rows, cols] tensor")
    if radius < 1:
        raise ValueError("radius must be at least 1")
    _c(type_layer)
    rows, cols = type_layer.shape
    if out is None:
        out = torch.empty((rows, cols), dtype=torch.uint8, device=type_layer.device)
    if count is None:
        count = torch.zeros(1, dtype=torch.int64, device=type_layer.device)
    if table is None:
        table = happy_threshold_table(homophily, radius)
    if torus and halo_top is None and halo_bot is None and rows > 0:
        if rows < 2 * radius + 1:
            raise ValueError("torus needs rows >= 2*radius+1")
        halo_top = type_layer[rows - radius:]
        halo_bot = type_layer[:radius]
    with torch.cuda.device(type_layer.device):
        N.check(N.lib().mb_schelling_happy_i8(
            N.ptr(type_layer), N.ptr(halo_top), N.ptr(halo_bot), N.ptr(_c(out)), rows, cols, int(radius),
            int(bool(torus)), table.ctypes.data, int(bool(binary_types)), N.ptr(count),
            N.stream_ptr(type_layer.device)))
    return out, count


class RelocationWorkspace:
    """Scratch of mb_schelling_relocate on one GPU: the 16-byte-per-cell slot table + the bitmap of the agents that stay
    put (both rebuilt from the layers at the start of every step) + the private mover / fallback scratch."""

    def __init__(self, type_layer: torch.Tensor, max_movers: int):
        import ctypes

        _need_cuda(type_layer)
        if type_layer.dtype != torch.int8 or type_layer.dim() != 2:
            raise ValueError("expected the int8 [rows, cols] type layer")
        self.rows, self.cols = type_layer.shape
        self.ncells, self.max_movers = int(type_layer.numel()), int(max_movers)
        nbytes = ctypes.c_size_t(0)
        N.check(N.lib().mb_reloc_workspace_bytes(self.max_movers, ctypes.byref(nbytes)))
        self.nbytes = int(nbytes.value)
        self.buf = torch.empty(max(self.nbytes, 1), dtype=torch.uint8, device=type_layer.device)
        tbytes = ctypes.c_size_t(0)
        N.check(N.lib().mb_reloc_table_bytes(self.rows, self.cols, ctypes.byref(tbytes)))
        self.slots_bytes = int(tbytes.value)     # 16-byte slots + the bitmap of the agents that stay put
        self.slots = torch.empty(max(-(-self.slots_bytes // 8), 1), dtype=torch.int64, device=type_layer.device)

    def sync(self, type_layer: torch.Tensor) -> None:
        """Kept for callers of the first version: the slot table is scratch now (rebuilt from the layers at the start of
        every relocation), so edits of the type layer need no notification."""
        if type_layer.dtype != torch.int8 or type_layer.numel() != self.ncells:
            raise ValueError("expected the int8 type layer this workspace was sized for")


def schelling_relocate(type_layer: torch.Tensor, happy: torch.Tensor, agent_id: torch.Tensor,
                       ws: RelocationWorkspace, seed: int, epoch: int,
                       agent_cell: torch.Tensor | None = None) -> tuple[int, int]:
    """shuffle_do("step") of the Schelling agents; returns (movers, fixed-point passes)."""
    import ctypes

    _need_cuda(type_layer, happy, agent_id, agent_cell)
    if type_layer.dtype != torch.int8 or happy.dtype != torch.uint8 or agent_id.dtype not in (torch.uint32, torch.int32):
        raise ValueError("schelling_relocate: expected int8 type, uint8 happy, uint32 agent_id layers")
    if ws.ncells != type_layer.numel():
        raise ValueError("workspace was sized for a different grid")
    movers, iters = ctypes.c_int64(0), ctypes.c_int(0)
    with torch.cuda.device(type_layer.device):
        N.check(N.lib().mb_schelling_relocate(
            N.ptr(_c(type_layer)), N.ptr(_c(happy)), N.ptr(_c(agent_id)), N.ptr(agent_cell), N.ptr(ws.slots),
            ws.slots_bytes, ws.rows, ws.cols, int(seed) & 0xFFFFFFFFFFFFFFFF, int(epoch) & 0xFFFFFFFF, N.ptr(ws.buf), ws.nbytes,
            ws.max_movers, ctypes.byref(movers), ctypes.byref(iters), N.stream_ptr(type_layer.device)))
    return int(movers.value), int(iters.value)


class SchellingSmallStep:
    """The whole Schelling model step of a small grid (<= 12800 cells, e.g. BASELINE config 1: 100 x 100) as ONE
    kernel launch with nothing read back (csrc/schelling_small.cu): shuffle_do("step") + do("assign_state").

    `out` is a device int64[4]: model.happy, movers, status, passes; `check()` reads it (synchronises) and raises the
    reference's ValueError for a completely full grid (grid.py:311-315)."""

    def __init__(self, type_layer: torch.Tensor, max_movers: int, homophily: float, radius: int, torus: bool = False):
        import ctypes

        _need_cuda(type_layer)
        if type_layer.dtype != torch.int8 or type_layer.dim() != 2:
            raise ValueError("expected the int8 [rows, cols] type layer")
        self.rows, self.cols = type_layer.shape
        if not self.supports(self.rows * self.cols):
            raise ValueError("SchellingSmallStep: the grid has more cells than the single-launch path holds")
        self.radius, self.torus = int(radius), bool(torus)
        self.max_movers = int(max_movers)
        self.table = happy_threshold_table(homophily, radius)
        nbytes = ctypes.c_size_t(0)
        N.check(N.lib().mb_schelling_small_workspace_bytes(self.max_movers, ctypes.byref(nbytes)))
        self.nbytes = int(nbytes.value)
        self.buf = torch.empty(max(self.nbytes, 1), dtype=torch.uint8, device=type_layer.device)
        self.out = torch.zeros(4, dtype=torch.int64, device=type_layer.device)
        self._bound = None

    def __getstate__(self):
        state = dict(self.__dict__)
        state["_bound"] = None          # (bound addresses and the ctypes entry point are per process)
        return state

    @staticmethod
    def supports(ncells: int) -> bool:
        return 0 < ncells <= N.lib().mb_schelling_small_max_cells()

    def step(self, type_layer, happy, agent_id, seed: int, epoch: int, agent_cell=None, empty_layer=None,
             relocate: bool = True, nsteps: int = 1, happy_current: bool = False) -> torch.Tensor:
        # A 100 x 100 step is ~10 us of kernel: the host side of the call must not cost more.  The buffers of a model do
        # not change between steps, so their addresses are bound once (keyed on the tensors' identity and storage) and a
        # step only passes (seed, epoch, relocate) and the current stream.
        key = (type_layer.data_ptr(), happy.data_ptr(), agent_id.data_ptr(), None if agent_cell is None else agent_cell.data_ptr(),
               None if empty_layer is None else empty_layer.data_ptr())
        bound = self._bound
        if bound is None or bound[0] != key:
            _need_cuda(type_layer, happy, agent_id, agent_cell, empty_layer)
            for t in (type_layer, happy, agent_id, agent_cell, empty_layer):
                if t is not None and not t.is_contiguous():
                    raise ValueError("SchellingSmallStep.step expects contiguous layers")
            dev = type_layer.device
            index = dev.index if dev.index is not None else torch.cuda.current_device()
            bound = self._bound = (key, index, N.lib().mb_schelling_step_small, self.table.ctypes.data, N.ptr(self.buf), N.ptr(self.out))
        _, index, fn, table, buf, out = bound
        seed, epoch, relocate = int(seed) & 0xFFFFFFFFFFFFFFFF, int(epoch) & 0xFFFFFFFF, int(bool(relocate))
        if torch.cuda.current_device() == index:
            rc = fn(key[0], key[1], key[2], key[3], key[4], self.rows, self.cols, self.radius, int(self.torus), table, seed, epoch,
                    relocate, int(nsteps), int(bool(happy_current)), buf, self.nbytes, self.max_movers, out,
                    torch.cuda.current_stream().cuda_stream)
        else:
            with torch.cuda.device(index):
                rc = fn(key[0], key[1], key[2], key[3], key[4], self.rows, self.cols, self.radius, int(self.torus), table, seed,
                        epoch, relocate, int(nsteps), int(bool(happy_current)), buf, self.nbytes, self.max_movers, out,
                    torch.cuda.current_stream().cuda_stream)
        if rc:
            N.check(rc)
        return self.out

    def check(self) -> tuple[int, int, int]:
        """(model.happy, movers, passes) of the last step; raises on a failed step.  Synchronises."""
        happy, movers, status, passes = (int(x) for x in self.out.tolist())
        if status == 1:
            raise ValueError("Grid is completely full. No empty cells available. Cannot select a random empty cell.")
        if status:
            raise N.MesaB200Error(f"mb_schelling_step_small failed with status {status}")
        return happy, movers, passes


# ----------------------------------------------------------------------------- sort
def sort_pairs(keys: torch.Tensor, vals: torch.Tensor, begin_bit: int = 0, end_bit: int | None = None):
    """Stable radix sort of (keys, vals); keys int32/int64 viewed as unsigned, vals int32.  Returns new tensors."""
    import ctypes

    _need_cuda(keys, vals)
    n = keys.numel()
    wide = keys.dtype in (torch.int64, torch.uint64)
    if end_bit is None:
        end_bit = 64 if wide else 32
    k0, v0 = keys.clone(), vals.clone()
    k1, v1 = torch.empty_like(k0), torch.empty_like(v0)
    nbytes = ctypes.c_size_t(0)
    N.check(N.lib().mb_sort_workspace_bytes(n, ctypes.byref(nbytes)))
    ws = torch.empty(nbytes.value, dtype=torch.uint8, device=keys.device)
    which = ctypes.c_int(0)
    fn = N.lib().mb_sort_pairs_u64 if wide else N.lib().mb_sort_pairs_u32
    with torch.cuda.device(keys.device):
        N.check(fn(N.ptr(k0), N.ptr(v0), N.ptr(k1), N.ptr(v1), n, begin_bit, end_bit, N.ptr(ws), nbytes.value,
                   ctypes.byref(which), N.stream_ptr(keys.device)))
    return (k1, v1) if which.value else (k0, v0)


def argsort_stable(keys: torch.Tensor, descending: bool = False) -> torch.Tensor:
    """Indices that sort a one-dimensional device column, equal keys in their original order (what `sorted(..., key=)`
    and `sorted(..., reverse=True)` give the reference, agentset.py:443-472): the column's bits are mapped to unsigned
    integers of the same order (sign flip for signed integers, sign-magnitude to offset for floats; complemented for a
    descending sort) and sorted with the stable radix sort of csrc/sort.cu.  int64 result."""
    if keys.dim() != 1:
        raise ValueError("argsort_stable expects a one-dimensional column")
    n = keys.numel()
    if n >= 2**31:
        raise ValueError("argsort_stable handles fewer than 2^31 rows")
    dt = keys.dtype
    if dt in (torch.bool, torch.uint8, torch.int8, torch.int16, torch.int32):
        u = keys.to(torch.int32) ^ (-2**31)                       # signed order -> unsigned order of the same bits
    elif dt == torch.int64:
        u = keys ^ (-2**63)
    elif dt in (torch.float16, torch.bfloat16, torch.float32):
        b = (keys.to(torch.float32) + 0.0).contiguous().view(torch.int32)    # + 0.0: -0.0 and 0.0 compare equal in Python
        u = torch.where(b < 0, ~b, b ^ (-2**31))                  # negative floats: all bits flipped, others: sign bit set
    elif dt == torch.float64:
        b = (keys + 0.0).contiguous().view(torch.int64)
        u = torch.where(b < 0, ~b, b ^ (-2**63))
    else:
        raise TypeError(f"argsort_stable: unsupported column dtype {dt}")
    if descending:
        u = ~u
    vals = torch.arange(n, dtype=torch.int32, device=keys.device)
    if n == 0:
        return vals.to(torch.int64)
    _, order = sort_pairs(u.contiguous(), vals)
    return order.to(torch.int64)


def nonzero_indices(mask: torch.Tensor) -> torch.Tensor:
    """Flat indices of the set entries of a bool / uint8 tensor in ascending order (np.argwhere / np.flatnonzero order,
    grid.py:308-316): one stable compaction of the index column (csrc/agents.cu).  int64 result."""
    flat = mask.reshape(-1)
    if flat.dtype not in (torch.bool, torch.uint8):
        flat = flat != 0
    n = flat.numel()
    if n == 0:
        return torch.empty(0, dtype=torch.int64, device=mask.device)
    plan = CompactionPlan(flat.contiguous())
    if n < 2**31:
        return plan.apply(torch.arange(n, dtype=torch.int32, device=mask.device)).to(torch.int64)
    return plan.apply(torch.arange(n, dtype=torch.int64, device=mask.device))


def offsets_from_counts(counts: torch.Tensor) -> torch.Tensor:
    """int64 [n + 1] row offsets of a CSR result from its int32 per-row counts (mb_offsets_from_counts)."""
    import ctypes

    _need_cuda(counts)
    if counts.dtype != torch.int32 or counts.dim() != 1:
        raise ValueError("offsets_from_counts expects an int32 [n] tensor")
    n = counts.numel()
    if n and int(layer_reduce(counts, "sum").item()) >= 2**32:
        raise ValueError("the query result holds 2^32 or more entries; query in smaller batches")
    offsets = torch.empty(n + 1, dtype=torch.int64, device=counts.device)
    words = ctypes.c_size_t(0)
    N.check(N.lib().mb_scan_scratch_words(n + 1, ctypes.byref(words)))
    scratch = torch.empty(n + 1 + int(words.value), dtype=torch.int32, device=counts.device)
    with torch.cuda.device(counts.device):
        N.check(N.lib().mb_offsets_from_counts(N.ptr(_c(counts)), n, N.ptr(offsets), N.ptr(scratch),
                                               N.stream_ptr(counts.device)))
    return offsets


# ----------------------------------------------------------------------------- continuous space
class CellListWorkspace:
    """Scratch of the cell-list kernels for up to `n` agents and query radius >= `radius`."""

    def __init__(self, n: int, size, radius: float, device):
        import ctypes

        nbytes = ctypes.c_size_t(0)
        N.check(N.lib().mb_celllist_workspace_bytes(int(n), float(size[0]), float(size[1]), float(radius),
                                                    ctypes.byref(nbytes)))
        self.n, self.radius, self.nbytes = int(n), float(radius), int(nbytes.value)
        self.buf = torch.empty(max(self.nbytes, 1), dtype=torch.uint8, device=device)


def boids_step(pos: torch.Tensor, direction: torch.Tensor, size, *, vision: float, separation: float,
               cohere: float = 0.03, separate: float = 0.015, match: float = 0.05, speed: float = 1.0,
               torus: bool = True, origin=(0.0, 0.0), ws: CellListWorkspace | None = None,
               pos_out: torch.Tensor | None = None, dir_out: torch.Tensor | None = None,
               neighbor_count: torch.Tensor | None = None):
    """One synchronous Boid.step of all agents (float32 [n,2] state).  Returns (pos_out, dir_out)."""
    _need_cuda(pos, direction, pos_out, dir_out, neighbor_count)
    if pos.dtype != torch.float32 or direction.dtype != torch.float32 or pos.shape != direction.shape or pos.dim() != 2 or pos.shape[1] != 2:
        raise ValueError("boids_step expects float32 [n, 2] positions and directions")
    n = pos.shape[0]
    if ws is None:
        ws = CellListWorkspace(n, size, vision, pos.device)
    if pos_out is None:
        pos_out = torch.empty_like(pos)
    if dir_out is None:
        dir_out = torch.empty_like(direction)
    with torch.cuda.device(pos.device):
        N.check(N.lib().mb_boids_step_f32(
            N.ptr(_c(pos)), N.ptr(_c(direction)), N.ptr(_c(pos_out)), N.ptr(_c(dir_out)), n, float(origin[0]),
            float(origin[1]), float(size[0]), float(size[1]), int(bool(torus)), float(vision), float(separation),
            float(cohere), float(separate), float(match), float(speed), N.ptr(neighbor_count), N.ptr(ws.buf),
            ws.nbytes, N.stream_ptr(pos.device)))
    return pos_out, dir_out


def activation_order(seed: int, epoch: int, n: int, device) -> torch.Tensor:
    """order[r] = index of the agent activated r-th in shuffle number `epoch` (ascending Philox priority64), int32."""
    import ctypes

    dev = torch.device(device)
    if dev.type != "cuda":
        raise RuntimeError("mesa_b200 runs on CUDA tensors only (no CPU fallback)")
    nbytes = ctypes.c_size_t(0)
    N.check(N.lib().mb_activation_order_workspace_bytes(int(n), ctypes.byref(nbytes)))
    ws = torch.empty(max(int(nbytes.value), 1), dtype=torch.uint8, device=dev)
    out = torch.empty(int(n), dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        N.check(N.lib().mb_activation_order(int(seed) & 0xFFFFFFFFFFFFFFFF, int(epoch) & 0xFFFFFFFF, int(n), N.ptr(out),
                                            N.ptr(ws), ws.numel(), N.stream_ptr(dev)))
    return out


class BoidsSeqWorkspace:
    """Scratch of the sequential Boids step for up to `n` agents (cell list with edge >= vision + reach, priorities,
    float64 intra-step state, done flags, pending lists)."""

    def __init__(self, n: int, size, vision: float, reach: float, device):
        import ctypes

        nbytes = ctypes.c_size_t(0)
        N.check(N.lib().mb_boids_seq_workspace_bytes(int(n), float(size[0]), float(size[1]), float(vision),
                                                     float(reach), ctypes.byref(nbytes)))
        self.n, self.vision, self.reach, self.nbytes = int(n), float(vision), float(reach), int(nbytes.value)
        self.buf = torch.empty(max(self.nbytes, 1), dtype=torch.uint8, device=device)


def boids_reach(direction: torch.Tensor, speed: float) -> float:
    """Upper bound on the distance a boid moves in one step: a boid with neighbours moves `speed` along a unit
    vector (agents.py:87-92), a lonely one `speed * |direction|` with its direction unchanged -- so the bound computed
    once from the initial directions holds for the whole run."""
    norm = float(torch.linalg.vector_norm(direction.to(torch.float64), dim=1).max().item()) if direction.numel() else 1.0
    return float(speed) * max(1.0, norm) * (1.0 + 1e-12)


def boids_step_sequential(pos: torch.Tensor, direction: torch.Tensor, size, *, seed: int, epoch: int, vision: float,
                          separation: float, cohere: float = 0.03, separate: float = 0.015, match: float = 0.05,
                          speed: float = 1.0, torus: bool = True, origin=(0.0, 0.0), reach: float | None = None,
                          ws: BoidsSeqWorkspace | None = None, pos_out: torch.Tensor | None = None,
                          dir_out: torch.Tensor | None = None, neighbor_count: torch.Tensor | None = None):
    """One Boid step of all agents in the reference's sequential shuffle_do order (ascending Philox priority of
    (seed, epoch, agent)).  Returns (pos_out, dir_out, passes)."""
    import ctypes

    _need_cuda(pos, direction, pos_out, dir_out, neighbor_count)
    if pos.dtype != torch.float32 or direction.dtype != torch.float32 or pos.shape != direction.shape or pos.dim() != 2 or pos.shape[1] != 2:
        raise ValueError("boids_step_sequential expects float32 [n, 2] positions and directions")
    n = pos.shape[0]
    if reach is None:
        reach = boids_reach(direction, speed)
    if ws is None:
        ws = BoidsSeqWorkspace(n, size, vision, reach, pos.device)
    if pos_out is None:
        pos_out = torch.empty_like(pos)
    if dir_out is None:
        dir_out = torch.empty_like(direction)
    passes = ctypes.c_int(0)
    with torch.cuda.device(pos.device):
        N.check(N.lib().mb_boids_step_seq_f32(
            N.ptr(_c(pos)), N.ptr(_c(direction)), N.ptr(_c(pos_out)), N.ptr(_c(dir_out)), n, float(origin[0]),
            float(origin[1]), float(size[0]), float(size[1]), int(bool(torus)), float(vision), float(separation),
            float(cohere), float(separate), float(match), float(speed), float(reach), int(seed) & 0xFFFFFFFFFFFFFFFF,
            int(epoch) & 0xFFFFFFFF, N.ptr(neighbor_count), N.ptr(ws.buf), ws.nbytes, ctypes.byref(passes),
            N.stream_ptr(pos.device)))
    return pos_out, dir_out, int(passes.value)


def boids_resort(pos: torch.Tensor, direction: torch.Tensor, size, *, vision: float, ids: torch.Tensor | None = None,
                 torus: bool = True, origin=(0.0, 0.0), ws: CellListWorkspace | None = None):
    """The population rewritten in cell-list order (mb_boids_resort_f32): returns (pos, direction, ids, perm) with
    row i = the boid at sorted slot i, perm[i] (int32) = its previous row, ids (int64) its unique id (`ids` given) or
    its previous row.  A model that re-sorts every few steps keeps the gathers / scatters of `boids_step` coalesced."""
    _need_cuda(pos, direction, ids)
    if pos.dtype != torch.float32 or direction.dtype != torch.float32 or pos.shape != direction.shape or pos.dim() != 2 or pos.shape[1] != 2:
        raise ValueError("boids_resort expects float32 [n, 2] positions and directions")
    n = pos.shape[0]
    if ids is not None and (ids.dtype != torch.int64 or tuple(ids.shape) != (n,)):
        raise ValueError("ids must be int64 [n]")
    if ws is None:
        ws = CellListWorkspace(n, size, vision, pos.device)
    pos_out, dir_out = torch.empty_like(pos), torch.empty_like(direction)
    ids_out = torch.empty(n, dtype=torch.int64, device=pos.device)
    perm = torch.empty(n, dtype=torch.int32, device=pos.device)
    with torch.cuda.device(pos.device):
        N.check(N.lib().mb_boids_resort_f32(
            N.ptr(_c(pos)), N.ptr(_c(direction)), N.ptr(None if ids is None else _c(ids)), N.ptr(pos_out), N.ptr(dir_out),
            N.ptr(ids_out), N.ptr(perm), n, float(origin[0]), float(origin[1]), float(size[0]), float(size[1]),
            int(bool(torus)), float(vision), N.ptr(ws.buf), ws.nbytes, N.stream_ptr(pos.device)))
    return pos_out, dir_out, ids_out, perm


def radius_query(pos: torch.Tensor, queries: torch.Tensor, size, radius: float, torus: bool = True,
                 origin=(0.0, 0.0), exclude: torch.Tensor | None = None, ws: CellListWorkspace | None = None):
    """Batched get_agents_in_radius: CSR (offsets[nq+1] int64, indices int32, distances float64)."""
    _need_cuda(pos, queries, exclude)
    n, nq = pos.shape[0], queries.shape[0]
    if ws is None:
        ws = CellListWorkspace(n, size, radius, pos.device)
    counts = torch.zeros(nq, dtype=torch.int32, device=pos.device)
    args = (float(origin[0]), float(origin[1]), float(size[0]), float(size[1]), int(bool(torus)))
    with torch.cuda.device(pos.device):
        st = N.stream_ptr(pos.device)
        N.check(N.lib().mb_radius_query_f32(N.ptr(_c(pos)), n, *args, N.ptr(_c(queries)), N.ptr(exclude), nq,
                                            float(radius), N.ptr(counts), None, None, None, N.ptr(ws.buf), ws.nbytes,
                                            0, st))
        offsets = offsets_from_counts(counts)
        total = int(offsets[-1].item())
        idx = torch.empty(total, dtype=torch.int32, device=pos.device)
        dist = torch.empty(total, dtype=torch.float64, device=pos.device)
        N.check(N.lib().mb_radius_query_f32(N.ptr(pos), n, *args, N.ptr(queries), N.ptr(exclude), nq, float(radius),
                                            None, N.ptr(offsets), N.ptr(idx), N.ptr(dist), N.ptr(ws.buf), ws.nbytes,
                                            1, st))
    return offsets, idx, dist


def knn_cell_edge(n: int, size, k: int) -> float:
    """Cell edge for a k-nearest search: about k + 1 agents per 3 x 3 block at uniform density."""
    area = float(size[0]) * float(size[1])
    edge = (area * (k + 1) / (9.0 * max(n, 1))) ** 0.5
    return min(max(edge, 1e-6 * max(float(size[0]), float(size[1]))), max(float(size[0]), float(size[1])))


def knn_query(pos: torch.Tensor, queries: torch.Tensor, size, k: int, torus: bool = True, origin=(0.0, 0.0),
              exclude: torch.Tensor | None = None, cell_edge: float | None = None):
    """Batched get_k_nearest_agents (continuous_space.py:249-283): (indices int32 [nq, k], distances float64 [nq, k]),
    rows ascending in (distance, storage index), padded with -1 / inf when fewer than k agents exist."""
    _need_cuda(pos, queries, exclude)
    n, nq = pos.shape[0], queries.shape[0]
    if k < 0:
        raise ValueError("k must be non-negative")
    idx = torch.empty((nq, k), dtype=torch.int32, device=pos.device)
    dist = torch.empty((nq, k), dtype=torch.float64, device=pos.device)
    if nq == 0 or k == 0:
        return idx, dist
    edge = float(cell_edge) if cell_edge is not None else knn_cell_edge(n, size, k)
    ws = CellListWorkspace(max(n, 1), size, edge, pos.device)
    with torch.cuda.device(pos.device):
        N.check(N.lib().mb_knn_query_f32(N.ptr(_c(pos)), n, float(origin[0]), float(origin[1]), float(size[0]),
                                         float(size[1]), int(bool(torus)), N.ptr(_c(queries)), N.ptr(exclude), nq, int(k),
                                         edge, N.ptr(idx), N.ptr(dist), N.ptr(ws.buf), ws.nbytes, 0,
                                         N.stream_ptr(pos.device)))
    return idx, dist


def point_distances(pos: torch.Tensor, point, size, torus: bool, sel: torch.Tensor | None = None,
                    distances: bool = True, deltas: bool = False):
    """calculate_distances / calculate_difference_vector of one point (continuous_space.py:169-235) against all
    agents or the agents `sel` (int32 indices): (float64 [m] or None, float64 [m, 2] or None)."""
    _need_cuda(pos, sel)
    m = pos.shape[0] if sel is None else sel.shape[0]
    d = torch.empty(m, dtype=torch.float64, device=pos.device) if distances else None
    v = torch.empty((m, 2), dtype=torch.float64, device=pos.device) if deltas else None
    with torch.cuda.device(pos.device):
        N.check(N.lib().mb_point_distances_f32(N.ptr(_c(pos)), N.ptr(sel), m, float(size[0]), float(size[1]),
                                               int(bool(torus)), float(point[0]), float(point[1]), N.ptr(d), N.ptr(v),
                                               N.stream_ptr(pos.device)))
    return d, v


def point_distances_nd(pos: torch.Tensor, point, size, torus: bool, sel: torch.Tensor | None = None,
                       distances: bool = True, deltas: bool = False):
    """point_distances for positions [n, ndim] of any dimension 1..8: (float64 [m] or None, float64 [m, ndim] or None)."""
    import ctypes

    _need_cuda(pos, sel)
    ndim = int(pos.shape[1])
    if len(point) != ndim or len(size) != ndim:
        raise ValueError(f"point / size must have {ndim} components")
    m = pos.shape[0] if sel is None else sel.shape[0]
    d = torch.empty(m, dtype=torch.float64, device=pos.device) if distances else None
    v = torch.empty((m, ndim), dtype=torch.float64, device=pos.device) if deltas else None
    arr = ctypes.c_double * ndim
    with torch.cuda.device(pos.device):
        N.check(N.lib().mb_point_distances_nd_f32(N.ptr(_c(pos)), ndim, N.ptr(sel), m, arr(*[float(s) for s in size]),
                                                  int(bool(torus)), arr(*[float(p) for p in point]), N.ptr(d), N.ptr(v),
                                                  N.stream_ptr(pos.device)))
    return d, v


# ----------------------------------------------------------------------------- Boltzmann wealth
class BoltzmannWorkspace:
    """Scratch + per-cell list order of mb_boltzmann_step for `n` agents (persistent between steps)."""

    def __init__(self, cell: torch.Tensor, rows: int, cols: int):
        import ctypes

        _need_cuda(cell)
        if cell.dtype != torch.int32:
            raise ValueError("cell must be int32 (uint32 bit pattern)")
        self.n, self.rows, self.cols = cell.numel(), int(rows), int(cols)
        nbytes = ctypes.c_size_t(0)
        N.check(N.lib().mb_boltzmann_workspace_bytes(self.n, ctypes.byref(nbytes)))
        self.nbytes = int(nbytes.value)
        self.buf = torch.empty(max(self.nbytes, 1), dtype=torch.uint8, device=cell.device)
        self.iterations = torch.zeros(1, dtype=torch.int32, device=cell.device)
        self.reset(cell)

    def reset(self, cell: torch.Tensor) -> None:
        """List order = registration order (after placing agents; model.py:70-74)."""
        with torch.cuda.device(cell.device):
            N.check(N.lib().mb_boltzmann_init(N.ptr(_c(cell)), self.n, self.rows, self.cols, N.ptr(self.buf),
                                              self.nbytes, N.stream_ptr(cell.device)))

    def check(self) -> None:
        """Raise if any step since `reset` could not be ordered exactly on the device.  Synchronises."""
        import ctypes

        status = ctypes.c_uint64(0)
        with torch.cuda.device(self.buf.device):
            N.check(N.lib().mb_boltzmann_status(N.ptr(self.buf), self.nbytes, self.n, ctypes.byref(status),
                                                N.stream_ptr(self.buf.device)))
        if status.value:
            raise N.MesaB200Error("boltzmann_step: more than 4096 cells held more than 32 agents in one step")


def boltzmann_step(cell: torch.Tensor, wealth: torch.Tensor, ws: BoltzmannWorkspace, seed: int, epoch: int,
                   torus: bool = False) -> torch.Tensor:
    """shuffle_do("step") of all MoneyAgents.  Only enqueues (no host synchronisation); returns the device int32[1]
    holding the number of give-resolution rounds (read it lazily).  `ws.check()` raises if a step was not exact."""
    _need_cuda(cell, wealth)
    if cell.dtype != torch.int32 or wealth.dtype != torch.int32 or cell.numel() != ws.n or wealth.numel() != ws.n:
        raise ValueError("boltzmann_step expects int32 cell / wealth tensors matching the workspace")
    with torch.cuda.device(cell.device):
        N.check(N.lib().mb_boltzmann_step(N.ptr(_c(cell)), N.ptr(_c(wealth)), ws.n, ws.rows, ws.cols,
                                          int(bool(torus)), int(seed) & 0xFFFFFFFFFFFFFFFF, int(epoch) & 0xFFFFFFFF,
                                          N.ptr(ws.buf), ws.nbytes, N.ptr(ws.iterations), N.stream_ptr(cell.device)))
    return ws.iterations


# ----------------------------------------------------------------------------- property layers
_DTYPES = {torch.uint8: 0, torch.bool: 0, torch.int32: 1, torch.float32: 2, torch.float64: 3}
_BINARY = {"add": 0, "sub": 1, "mul": 2, "min": 3, "max": 4}
_REDUCE = {"sum": 0, "min": 1, "max": 2, "count_nonzero": 3}


def _dt(t: torch.Tensor) -> int:
    try:
        return _DTYPES[t.dtype]
    except KeyError:
        raise TypeError(f"property-layer kernels support uint8/bool/int32/float32/float64, not {t.dtype}") from None


def layer_fill(layer: torch.Tensor, value) -> torch.Tensor:
    """layer[:] = value  (np.full of grid.py:146 on an existing buffer)."""
    _need_cuda(layer)
    with torch.cuda.device(layer.device):
        N.check(N.lib().mb_layer_fill(N.ptr(_c(layer)), layer.numel(), _dt(layer), float(value),
                                      N.stream_ptr(layer.device)))
    return layer


def layer_affine_clamp(layer: torch.Tensor, alpha=1.0, beta=0.0, lo=float("-inf"), hi=float("inf"),
                       out: torch.Tensor | None = None) -> torch.Tensor:
    """out = clip(alpha * layer + beta, lo, hi); out defaults to in place."""
    _need_cuda(layer, out)
    out = layer if out is None else out
    with torch.cuda.device(layer.device):
        N.check(N.lib().mb_layer_affine_clamp(N.ptr(_c(layer)), N.ptr(_c(out)), layer.numel(), _dt(layer),
                                              float(alpha), float(beta), float(lo), float(hi),
                                              N.stream_ptr(layer.device)))
    return out


def layer_binary(a: torch.Tensor, b: torch.Tensor, op: str, out: torch.Tensor | None = None) -> torch.Tensor:
    """out = a <op> b elementwise, op in add/sub/mul/min/max (e.g. np.minimum(sugar + 1, cap))."""
    _need_cuda(a, b, out)
    if a.shape != b.shape or a.dtype != b.dtype:
        raise ValueError("layer_binary: shape / dtype mismatch")
    out = torch.empty_like(a) if out is None else out
    with torch.cuda.device(a.device):
        N.check(N.lib().mb_layer_binary(N.ptr(_c(a)), N.ptr(_c(b)), N.ptr(_c(out)), a.numel(), _dt(a), _BINARY[op],
                                        N.stream_ptr(a.device)))
    return out


def layer_reduce(layer: torch.Tensor, op: str = "sum") -> torch.Tensor:
    """Aggregate a layer on the device; returns a 1-element tensor (int64 or float64)."""
    _need_cuda(layer)
    code = _REDUCE[op]
    as_int = layer.dtype in (torch.uint8, torch.bool, torch.int32) or op == "count_nonzero"
    res = torch.zeros(1, dtype=torch.int64 if as_int else torch.float64, device=layer.device)
    ws = torch.empty(1024, dtype=torch.float64, device=layer.device)
    with torch.cuda.device(layer.device):
        N.check(N.lib().mb_layer_reduce(N.ptr(_c(layer)), layer.numel(), _dt(layer), code, N.ptr(res), N.ptr(ws),
                                        ws.numel() * 8, N.stream_ptr(layer.device)))
    return res


def neighborhood_sum(layer: torch.Tensor, radius: int = 1, moore: bool = True, include_center: bool = False,
                     torus: bool = False, halo_top: torch.Tensor | None = None,
                     halo_bot: torch.Tensor | None = None, out: torch.Tensor | None = None) -> torch.Tensor:
    """Sum of `layer` over each cell's radius-r neighbourhood (cell.py:225-276) for all cells at once
    (int32 for integer layers, float64 for float layers)."""
    _need_cuda(layer, halo_top, halo_bot, out)
    if radius < 1:
        raise ValueError("radius must be at least 1")
    if layer.dim() != 2:
        if halo_top is not None or halo_bot is not None:
            raise ValueError("halo rows are a feature of two-dimensional (row-sharded) layers")
        return _neighborhood_sum_nd(layer, radius, moore, include_center, torus, out)
    rows, cols = layer.shape
    integer = layer.dtype in (torch.uint8, torch.bool, torch.int32)
    want = torch.int32 if integer else torch.float64
    if out is None:
        out = torch.empty((rows, cols), dtype=want, device=layer.device)
    elif out.shape != layer.shape or out.dtype != want or not out.is_contiguous():
        raise ValueError(f"neighborhood_sum: out must be a contiguous {want} tensor of the layer's shape")
    if torus and halo_top is None and halo_bot is None and rows > 0:
        if rows < 2 * radius + 1:
            raise ValueError("torus needs rows >= 2*radius+1")
        halo_top, halo_bot = layer[rows - radius:], layer[:radius]
    with torch.cuda.device(layer.device):
        N.check(N.lib().mb_neighborhood_sum(N.ptr(_c(layer)), N.ptr(halo_top), N.ptr(halo_bot), N.ptr(out), rows, cols,
                                            _dt(layer), int(radius), int(not moore), int(bool(include_center)),
                                            int(bool(torus)), N.stream_ptr(layer.device)))
    return out


def _neighborhood_sum_nd(layer, radius, moore, include_center, torus, out):
    """n-dimensional grids (grid.py:457-462, 488-501): 1 to 4 dimensions, one thread per cell."""
    import ctypes

    if not 1 <= layer.dim() <= 4:
        raise ValueError("neighborhood_sum supports layers of 1 to 4 dimensions")
    integer = layer.dtype in (torch.uint8, torch.bool, torch.int32)
    want = torch.int32 if integer else torch.float64
    if out is None:
        out = torch.empty(layer.shape, dtype=want, device=layer.device)
    elif out.shape != layer.shape or out.dtype != want or not out.is_contiguous():
        raise ValueError(f"neighborhood_sum: out must be a contiguous {want} tensor of the layer's shape")
    if torus and any(d < 2 * radius + 1 for d in layer.shape):
        raise ValueError("torus needs every dimension >= 2*radius+1")
    dims = (ctypes.c_int64 * layer.dim())(*layer.shape)
    with torch.cuda.device(layer.device):
        N.check(N.lib().mb_neighborhood_sum_nd(N.ptr(_c(layer)), N.ptr(out), layer.dim(), dims, _dt(layer), int(radius),
                                               int(not moore), int(bool(include_center)), int(bool(torus)),
                                               N.stream_ptr(layer.device)))
    return out


def layer_diffuse(layer: torch.Tensor, rate: float, torus: bool = False, out: torch.Tensor | None = None,
                  halo_top: torch.Tensor | None = None, halo_bot: torch.Tensor | None = None) -> torch.Tensor:
    """One diffusion step (each cell shares `rate` of its value equally among its 8 neighbours)."""
    _need_cuda(layer, out, halo_top, halo_bot)
    if layer.dim() != 2 or layer.dtype not in (torch.float32, torch.float64):
        raise ValueError("layer_diffuse expects a float32/float64 [rows, cols] layer")
    rows, cols = layer.shape
    out = torch.empty_like(layer) if out is None else out
    if torus and halo_top is None and halo_bot is None and rows > 0:
        if rows < 3:
            raise ValueError("torus needs rows >= 3")
        halo_top, halo_bot = layer[rows - 1], layer[0]
    with torch.cuda.device(layer.device):
        N.check(N.lib().mb_layer_diffuse(N.ptr(_c(layer)), N.ptr(halo_top), N.ptr(halo_bot), N.ptr(_c(out)), rows, cols,
                                         _dt(layer), float(rate), int(bool(torus)), N.stream_ptr(layer.device)))
    return out


# ----------------------------------------------------------------------------- dynamic populations
class CompactionPlan:
    """Keep flags + their exclusive prefix sum for one stable compaction shared by all columns of a
    population (csrc/agents.cu).  `kept` reads the survivor count back (one 4-byte copy)."""

    def __init__(self, keep: torch.Tensor):
        import ctypes

        _need_cuda(keep)
        if keep.dim() != 1:
            raise ValueError("keep flags must be one-dimensional")
        if keep.dtype == torch.bool:
            keep = keep.view(torch.uint8)
        elif keep.dtype != torch.uint8:
            raise TypeError("keep flags must be bool or uint8")
        self.keep = _c(keep)
        self.n = keep.numel()
        self.offsets = torch.empty(self.n + 1, dtype=torch.int32, device=keep.device)
        words = ctypes.c_size_t(0)
        N.check(N.lib().mb_scan_scratch_words(self.n + 1, ctypes.byref(words)))
        scratch = torch.empty(words.value, dtype=torch.int32, device=keep.device)
        with torch.cuda.device(keep.device):
            N.check(N.lib().mb_compact_plan(N.ptr(self.keep), self.n, N.ptr(self.offsets), N.ptr(scratch),
                                            N.stream_ptr(keep.device)))
        self._kept = None

    @property
    def kept(self) -> int:
        if self._kept is None:
            self._kept = int(self.offsets[self.n].item()) & 0xFFFFFFFF
        return self._kept

    def apply(self, column: torch.Tensor) -> torch.Tensor:
        """The surviving rows of `column` ([n, ...]) in their original order, as a new tensor."""
        _need_cuda(column)
        if column.shape[0] != self.n:
            raise ValueError(f"column has {column.shape[0]} rows, the plan {self.n}")
        col = _c(column)
        out = torch.empty((self.kept, *col.shape[1:]), dtype=col.dtype, device=col.device)
        row_bytes = col.element_size()
        for d in col.shape[1:]:
            row_bytes *= int(d)
        if self.n and row_bytes and self.kept:
            with torch.cuda.device(col.device):
                N.check(N.lib().mb_compact_rows(N.ptr(col), N.ptr(out), self.n, row_bytes, N.ptr(self.keep),
                                                N.ptr(self.offsets), N.stream_ptr(col.device)))
        return out
