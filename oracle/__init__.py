"""CPU oracle for the mesa_b200 hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
may import this package; nothing under mesa_b200/ does.  It holds

* mesa_oracle.c  -- plain-C restatement of the reference's algorithms (ctypes, below);
* philox.py      -- numpy Philox4x32-10 + the keyed draw protocol;
* replay.py      -- ReplayRandom: drives the UNMODIFIED reference with those draws;
* make_golden.py -- runs the live reference (build container only) and writes tests/golden/.

Parity pin: tests/test_oracle_golden.py checks every function here against fixtures that
make_golden.py produced from the live reference; tests/test_live_reference.py re-runs the
comparison LIVE (reference model vs this oracle, from fresh seeds) wherever /root/reference exists.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from ctypes import c_double, c_int, c_int64, c_uint32, c_uint64, c_void_p

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_build", "liboracle.so")

_lib = None


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "mesa_oracle.c")
    if force or not os.path.exists(LIB_PATH) or os.path.getmtime(src) > os.path.getmtime(LIB_PATH):
        subprocess.run(["make", "-C", _HERE, "-B" if force else "-s"], check=True, capture_output=True)
    return LIB_PATH


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(LIB_PATH)
    return _lib


def _p(a: np.ndarray):
    return a.ctypes.data_as(c_void_p)


def set_threads(n: int) -> None:
    lib().orc_set_threads(c_int(int(n)))


def philox4x32_10(ctr, key) -> np.ndarray:
    ctr = np.ascontiguousarray(ctr, dtype=np.uint32)
    key = np.ascontiguousarray(key, dtype=np.uint32)
    out = np.empty(4, dtype=np.uint32)
    lib().orc_philox4x32_10(_p(ctr), _p(key), _p(out))
    return out


def gol_step(state: np.ndarray, torus: bool = True) -> np.ndarray:
    """One synchronous Game of Life generation of a uint8 [rows, cols] layer."""
    state = np.ascontiguousarray(state, dtype=np.uint8)
    out = np.empty_like(state)
    rc = lib().orc_gol_step_u8(_p(state), _p(out), c_int64(state.shape[0]), c_int64(state.shape[1]),
                               c_int(bool(torus)))
    if rc != 0:
        raise ValueError("gol_step: torus needs both dimensions >= 3")
    return out


# ----------------------------------------------------------------------------- draw protocol
def priority64(seed: int, epoch: int, agent: int) -> int:
    f = lib().orc_priority64
    f.restype = c_uint64
    return int(f(c_uint64(seed), c_uint32(epoch), c_uint64(agent)))


def activation_order(seed: int, epoch: int, n: int) -> np.ndarray:
    order = np.empty(n, dtype=np.int64)
    rc = lib().orc_activation_order(c_uint64(seed), c_uint32(epoch), c_int64(n), _p(order))
    assert rc == 0
    return order


# ----------------------------------------------------------------------------- Schelling
class SchellingState:
    """Reference-shaped Schelling state: per-agent (registration order) cell/type/happy + occupancy."""

    def __init__(self, rows: int, cols: int, cell, atype, happy=None):
        self.rows, self.cols = int(rows), int(cols)
        self.cell = np.ascontiguousarray(cell, dtype=np.int64).copy()
        self.atype = np.ascontiguousarray(atype, dtype=np.int8).copy()
        n = len(self.cell)
        self.happy = np.zeros(n, np.uint8) if happy is None else np.ascontiguousarray(happy, dtype=np.uint8).copy()
        self.occ = np.full(self.rows * self.cols, -1, dtype=np.int64)
        self.occ[self.cell] = np.arange(n)
        assert (self.occ >= 0).sum() == n, "capacity-1 grid: two agents share a cell"

    @classmethod
    def from_type_grid(cls, grid: np.ndarray):
        """Agents registered row-major over occupied cells (schelling/model.py:72-81)."""
        grid = np.asarray(grid)
        flat = np.flatnonzero(grid.reshape(-1) >= 0)
        return cls(grid.shape[0], grid.shape[1], flat, grid.reshape(-1)[flat])

    @property
    def n(self) -> int:
        return len(self.cell)

    def type_grid(self) -> np.ndarray:
        g = np.full(self.rows * self.cols, -1, dtype=np.int8)
        g[self.cell] = self.atype
        return g.reshape(self.rows, self.cols)

    def happy_grid(self) -> np.ndarray:
        g = np.zeros(self.rows * self.cols, dtype=np.uint8)
        g[self.cell] = self.happy
        return g.reshape(self.rows, self.cols)

    def id_grid(self) -> np.ndarray:
        return self.occ.reshape(self.rows, self.cols).copy()

    def assign_state(self, homophily: float, radius: int = 1, torus: bool = False) -> int:
        f = lib().orc_schelling_assign
        f.restype = c_int64
        return int(f(c_int64(self.rows), c_int64(self.cols), c_int(radius), c_int(bool(torus)),
                     c_double(homophily), c_int64(self.n), _p(self.cell), _p(self.atype), _p(self.happy),
                     _p(self.occ)))

    def move_unhappy(self, seed: int, epoch: int) -> tuple[int, int]:
        """shuffle_do("step"); returns (movers, argwhere fallbacks)."""
        f = lib().orc_schelling_move
        f.restype = c_int64
        nfb = c_int64(0)
        movers = int(f(c_int64(self.rows * self.cols), c_int64(self.n), _p(self.cell), _p(self.happy),
                       _p(self.occ), c_uint64(seed), c_uint32(epoch), ctypes.byref(nfb)))
        if movers == -1:
            raise ValueError("Grid is completely full. No empty cells available. Cannot select a random empty cell.")
        assert movers >= 0
        return movers, int(nfb.value)


def happy_threshold_table(homophily: float, radius: int) -> np.ndarray:
    """min_similar[valid] with the reference's float arithmetic (agents.py:31-39)."""
    nmax = (2 * radius + 1) ** 2 - 1
    tbl = np.zeros(nmax + 1, dtype=np.uint8)
    for valid in range(nmax + 1):
        t = valid + 1  # never happy
        for similar in range(valid + 1):
            frac = similar / valid if valid > 0 else 0.0
            if not (frac < homophily):
                t = similar
                break
        tbl[valid] = t
    return tbl


# ----------------------------------------------------------------------------- continuous space / Boids
def boids_step(pos, direction, size, *, vision, separation, cohere=0.03, separate=0.015, match=0.05, speed=1.0,
               torus=True, origin=(0.0, 0.0), order=None):
    """One BoidFlockers step.  order=None -> synchronous (frozen snapshot); order=array -> the reference's
    sequential shuffle_do in that activation order.  Returns (pos, dir, neighbour counts), float64."""
    pos = np.ascontiguousarray(pos, dtype=np.float64)
    direction = np.ascontiguousarray(direction, dtype=np.float64)
    n = pos.shape[0]
    po, do = np.empty_like(pos), np.empty_like(direction)
    cnt = np.zeros(n, dtype=np.int64)
    od = None if order is None else np.ascontiguousarray(order, dtype=np.int64)
    rc = lib().orc_boids_step(c_int64(n), _p(pos), _p(direction), _p(po), _p(do), c_double(origin[0]),
                              c_double(origin[1]), c_double(size[0]), c_double(size[1]), c_int(bool(torus)),
                              c_double(vision), c_double(separation), c_double(cohere), c_double(separate),
                              c_double(match), c_double(speed), None if od is None else _p(od), _p(cnt))
    assert rc == 0
    return po, do, cnt


def radius_query(pos, point, size, radius, torus=True, exclude=-1):
    """(indices ascending, distances) of agents within `radius` (inclusive) of `point`."""
    pos = np.ascontiguousarray(pos, dtype=np.float64)
    n = pos.shape[0]
    idx = np.empty(n, dtype=np.int64)
    dist = np.empty(n, dtype=np.float64)
    f = lib().orc_radius_query
    f.restype = c_int64
    k = int(f(c_int64(n), _p(pos), c_double(point[0]), c_double(point[1]), c_double(size[0]), c_double(size[1]),
              c_int(bool(torus)), c_double(radius), c_int64(exclude), _p(idx), _p(dist)))
    return idx[:k].copy(), dist[:k].copy()


def point_distances(pos, point, size, torus=True):
    """calculate_distances (continuous_space.py:202-235): torus -> abs, min(d, size - d), squares summed axis by
    axis, sqrt; bounded -> Euclidean (scipy cdist there; the same sum of squares)."""
    pos = np.asarray(pos, dtype=np.float64)
    delta = np.abs(np.asarray(point, dtype=np.float64)[np.newaxis, :] - pos)
    if torus:
        delta = np.minimum(delta, np.asarray(size, dtype=np.float64) - delta)
    d2 = np.zeros(pos.shape[0])
    for axis in range(pos.shape[1]):
        d2 += delta[:, axis] ** 2
    return np.sqrt(d2)


def difference_vectors(pos, point, size, torus=True):
    """calculate_difference_vector (continuous_space.py:169-200)."""
    delta = np.asarray(pos, dtype=np.float64) - np.asarray(point, dtype=np.float64)[np.newaxis, :]
    if torus:
        inverse = delta - np.sign(delta) * np.asarray(size, dtype=np.float64)
        delta = np.where(np.abs(delta) < np.abs(inverse), delta, inverse)
    return delta


def knn_query(pos, point, size, k, torus=True, exclude=-1):
    """get_k_nearest_agents (continuous_space.py:249-283) with the documented tie rule made explicit: ascending
    (distance, storage index).  Returns (indices, distances) of min(k, n) agents."""
    d = point_distances(pos, point, size, torus)
    ids = np.arange(len(d))
    if exclude >= 0:
        keep = ids != exclude
        d, ids = d[keep], ids[keep]
    o = np.lexsort((ids, d))[:k]
    return ids[o], d[o]


# ----------------------------------------------------------------------------- Boltzmann wealth
class BoltzmannState:
    """Reference-shaped Boltzmann state: per-agent cell/wealth + per-cell agent lists in list order."""

    def __init__(self, rows: int, cols: int, cell, wealth=None, torus: bool = False):
        self.rows, self.cols, self.torus = int(rows), int(cols), bool(torus)
        self.cell = np.ascontiguousarray(cell, dtype=np.int64).copy()
        n = len(self.cell)
        self.wealth = (np.ones(n, np.int32) if wealth is None else np.ascontiguousarray(wealth, dtype=np.int32).copy())
        nc = self.rows * self.cols
        self.next, self.prev = np.empty(n, np.int64), np.empty(n, np.int64)
        self.head, self.tail = np.empty(nc, np.int64), np.empty(nc, np.int64)
        lib().orc_boltzmann_init_lists(c_int64(nc), c_int64(n), _p(self.cell), _p(self.next), _p(self.prev),
                                       _p(self.head), _p(self.tail))

    @property
    def n(self) -> int:
        return len(self.cell)

    def step(self, seed: int, epoch: int) -> None:
        rc = lib().orc_boltzmann_step(c_int64(self.rows), c_int64(self.cols), c_int(self.torus), c_int64(self.n),
                                      _p(self.cell), _p(self.wealth), _p(self.next), _p(self.prev), _p(self.head),
                                      _p(self.tail), c_uint64(seed), c_uint32(epoch))
        if rc == -1:
            raise LookupError("Cannot select random cell from empty collection")
        assert rc == 0

    def cell_lists(self) -> dict:
        """{cell: [agents in list order]} for non-empty cells."""
        out = {}
        for c in np.flatnonzero(self.head >= 0):
            lst, b = [], self.head[c]
            while b >= 0:
                lst.append(int(b))
                b = self.next[b]
            out[int(c)] = lst
        return out
