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
