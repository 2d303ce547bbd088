"""numpy restatements of the property-layer / neighbourhood kernels -- TEST INFRASTRUCTURE.

* neighborhood_sum follows Cell._neighborhood (mesa/discrete_space/cell.py:225-276): Moore radius r
  = Chebyshev square, von Neumann = Manhattan diamond (tests/discrete_space/test_discrete_space.py:380-398),
  clamped edges drop cells, torus wraps (grid.py:390-399); pinned against the live reference's
  neighbourhood sets in tests/golden/neighborhoods.npz.
* diffuse: PARITY UNPINNED -- the reference has no diffusion; this restates the NetLogo rule the
  device kernel documents (csrc/layers.cu).
"""
from __future__ import annotations

import numpy as np


def neighborhood_offsets(radius: int, moore: bool, include_center: bool):
    offs = []
    for di in range(-radius, radius + 1):
        for dj in range(-radius, radius + 1):
            if not moore and abs(di) + abs(dj) > radius:
                continue
            if di == 0 and dj == 0 and not include_center:
                continue
            offs.append((di, dj))
    return offs


def neighborhood_sum(layer: np.ndarray, radius=1, moore=True, include_center=False, torus=False) -> np.ndarray:
    layer = np.asarray(layer)
    rows, cols = layer.shape
    acc_t = np.float64 if layer.dtype.kind == "f" else np.int64
    out = np.zeros((rows, cols), dtype=acc_t)
    ii, jj = np.meshgrid(np.arange(rows), np.arange(cols), indexing="ij")
    for di, dj in neighborhood_offsets(radius, moore, include_center):
        ni, nj = ii + di, jj + dj
        if torus:
            out += layer[ni % rows, nj % cols].astype(acc_t)
        else:
            ok = (ni >= 0) & (ni < rows) & (nj >= 0) & (nj < cols)
            out[ok] += layer[ni[ok], nj[ok]].astype(acc_t)
    return out


def neighborhood_sum_nd(layer: np.ndarray, radius=1, moore=True, include_center=False, torus=False) -> np.ndarray:
    """The same for layers of any dimensionality (grid.py:457-462: Moore connections = product([-1,0,1]^n);
    :488-501: von Neumann = +-1 per axis; radius r by BFS = Chebyshev cube / Manhattan ball)."""
    from itertools import product

    layer = np.asarray(layer)
    acc_t = np.float64 if layer.dtype.kind == "f" else np.int64
    out = np.zeros(layer.shape, dtype=acc_t)
    grids = np.meshgrid(*(np.arange(d) for d in layer.shape), indexing="ij")
    for off in product(range(-radius, radius + 1), repeat=layer.ndim):
        if not moore and sum(abs(o) for o in off) > radius:
            continue
        if not any(off) and not include_center:
            continue
        idx = [g + o for g, o in zip(grids, off)]
        if torus:
            out += layer[tuple(i % d for i, d in zip(idx, layer.shape))].astype(acc_t)
        else:
            ok = np.ones(layer.shape, dtype=bool)
            for i, d in zip(idx, layer.shape):
                ok &= (i >= 0) & (i < d)
            out[ok] += layer[tuple(i[ok] for i in idx)].astype(acc_t)
    return out


def diffuse(layer: np.ndarray, rate: float, torus=False) -> np.ndarray:
    layer = np.asarray(layer)
    rows, cols = layer.shape
    x = layer.astype(np.float64)
    acc = np.zeros_like(x)
    k = np.zeros((rows, cols), dtype=np.int64)
    ii, jj = np.meshgrid(np.arange(rows), np.arange(cols), indexing="ij")
    for di in (-1, 0, 1):          # same accumulation order as the kernel
        for dj in (-1, 0, 1):
            if di == 0 and dj == 0:
                continue
            ni, nj = ii + di, jj + dj
            if torus:
                acc += x[ni % rows, nj % cols]
                k += 1
            else:
                ok = (ni >= 0) & (ni < rows) & (nj >= 0) & (nj < cols)
                acc[ok] += x[ni[ok], nj[ok]]
                k += ok
    return (x * (1.0 - rate * k / 8.0) + rate / 8.0 * acc).astype(layer.dtype)
