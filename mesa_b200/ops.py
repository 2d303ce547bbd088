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
    """Persistent scratch of mb_schelling_relocate for a grid of `ncells` cells."""

    def __init__(self, ncells: int, max_movers: int, device):
        import ctypes

        self.ncells, self.max_movers = int(ncells), int(max_movers)
        nbytes = ctypes.c_size_t(0)
        N.check(N.lib().mb_relocate_workspace_bytes(self.ncells, self.max_movers, ctypes.byref(nbytes)))
        self.nbytes = int(nbytes.value)
        self.buf = torch.empty(max(self.nbytes, 1), dtype=torch.uint8, device=device)
        with torch.cuda.device(device):
            N.check(N.lib().mb_relocate_workspace_init(N.ptr(self.buf), self.nbytes, self.ncells,
                                                       self.max_movers, N.stream_ptr(device)))


def schelling_relocate(type_layer: torch.Tensor, happy: torch.Tensor, agent_id: torch.Tensor,
                       ws: RelocationWorkspace, seed: int, epoch: int,
                       agent_cell: torch.Tensor | None = None) -> tuple[int, int]:
    """shuffle_do("step") of the Schelling agents; returns (movers, fixed-point iterations)."""
    import ctypes

    _need_cuda(type_layer, happy, agent_id, agent_cell)
    if type_layer.dtype != torch.int8 or happy.dtype != torch.uint8 or agent_id.dtype not in (torch.uint32, torch.int32):
        raise ValueError("schelling_relocate: expected int8 type, uint8 happy, uint32 agent_id layers")
    ncells = type_layer.numel()
    if ws.ncells != ncells:
        raise ValueError("workspace was sized for a different grid")
    movers, iters = ctypes.c_int64(0), ctypes.c_int(0)
    with torch.cuda.device(type_layer.device):
        N.check(N.lib().mb_schelling_relocate(
            N.ptr(_c(type_layer)), N.ptr(_c(happy)), N.ptr(_c(agent_id)), N.ptr(agent_cell), ncells,
            int(seed) & 0xFFFFFFFFFFFFFFFF, int(epoch) & 0xFFFFFFFF, N.ptr(ws.buf), ws.nbytes, ws.max_movers,
            ctypes.byref(movers), ctypes.byref(iters), N.stream_ptr(type_layer.device)))
    return int(movers.value), int(iters.value)
