"""DeviceContinuousSpace: device-resident mirror of mesa.experimental.continuous_space.ContinuousSpace.

Mirrors mesa/experimental/continuous_space/continuous_space.py:48-298: `dimensions`, `size`, `center`,
`torus`, `agent_positions`, `in_bounds`, `torus_correct`, `calculate_distances`,
`calculate_difference_vector`, `get_agents_in_radius`, `get_k_nearest_agents` -- same argument meaning,
same inclusive radius (`dist <= radius`), same errors.  Positions live in ONE float32 [n, 2] device
tensor (storage-index order = the reference's `_agent_positions` rows); agents are identified by
that index.  Distances are evaluated in float64 with the reference's operation order, so for
fp32-representable positions they are bit-identical to the reference's (csrc/continuous.cu).

Single-point queries launch one small kernel each; the batched `radius_query` /
`neighbors_in_radius` are the hot-path entry points (one cell-list build + one query kernel).

Spaces of 1 to 8 dimensions are supported like the reference's (capacity, ndims) storage (continuous_space.py:85-101):
outside two dimensions the point queries evaluate all n distances in one launch (`mb_point_distances_nd_f32`, what the
reference does in numpy) and the cell-list based batched queries are not available.
"""
from __future__ import annotations

import warnings
from random import Random

import numpy as np
import torch

from . import ops
from .agentset import DeviceAgent


class DeviceContinuousSpace:
    def __init__(self, dimensions, torus: bool = False, random: Random | None = None, n_agents: int = 100,
                 device="cuda"):
        if random is None:
            warnings.warn("Random number generator not specified, this can make models non-reproducible. "
                          "Please pass a random number generator explicitly", UserWarning, stacklevel=2)
            random = Random()
        self.random = random
        self.dimensions = np.asanyarray(dimensions, dtype=float)
        self.ndims = int(self.dimensions.shape[0])
        if self.dimensions.ndim != 2 or self.dimensions.shape[1] != 2 or not 1 <= self.ndims <= 8:
            raise ValueError("dimensions must be [[lo, hi], ...] with 1 to 8 axes")
        self.size = self.dimensions[:, 1] - self.dimensions[:, 0]
        self.center = np.sum(self.dimensions, axis=1) / 2
        self.torus = bool(torus)
        self.device = torch.device(device)
        self._agent_positions = torch.empty((int(n_agents), self.ndims), dtype=torch.float32, device=self.device)
        self._n_agents = 0
        self._ws = None

    # -- storage (continuous_space.py:108-167) ---------------------------------
    @property
    def agent_positions(self) -> torch.Tensor:
        return self._agent_positions[: self._n_agents]

    @property
    def width(self):
        return self.size[0]

    @property
    def height(self):
        return self.size[1]

    def add_agents(self, positions) -> torch.Tensor:
        """Append agents at `positions` ([k, ndims]); returns their storage indices.  Out-of-bounds points wrap on
        a torus and raise ValueError otherwise (continuous_space_agents.py:36-44)."""
        pos = torch.as_tensor(np.asarray(positions, dtype=np.float64).reshape(-1, self.ndims))
        lo = torch.as_tensor(self.dimensions[:, 0])
        hi = torch.as_tensor(self.dimensions[:, 1])
        bad = ~((pos >= lo) & (pos <= hi)).all(dim=1)
        if bad.any():
            if not self.torus:
                raise ValueError(f"point {pos[bad][0].tolist()} is outside the bounds of the space")
            size = torch.as_tensor(self.size)
            pos[bad] = lo + torch.remainder(pos[bad] - lo, size)
        k = pos.shape[0]
        need = self._n_agents + k
        cap = self._agent_positions.shape[0]
        if need > cap:
            # growth +max(20 %, 20) rows like continuous_space.py:121-129, at least what is needed
            new_cap = max(need, cap + max(int(cap * 0.2), 20))
            grown = torch.empty((new_cap, self.ndims), dtype=torch.float32, device=self.device)
            grown[: self._n_agents] = self._agent_positions[: self._n_agents]
            self._agent_positions = grown
        self._agent_positions[self._n_agents:need] = pos.to(torch.float32).to(self.device)
        idx = torch.arange(self._n_agents, need, device=self.device)
        self._n_agents = need
        self._ws = None
        return idx

    def remove_agent(self, index: int) -> int | None:
        """Swap-with-last removal (continuous_space.py:141-167); returns the index that moved into `index`."""
        last = self._n_agents - 1
        if not 0 <= index <= last:
            raise IndexError(index)
        moved = None
        if index != last:
            self._agent_positions[index] = self._agent_positions[last]
            moved = last
        self._n_agents -= 1
        self._ws = None
        return moved

    # -- geometry ---------------------------------------------------------------
    def in_bounds(self, point) -> bool:
        p = np.asanyarray(point)
        return bool(((p >= self.dimensions[:, 0]) & (p <= self.dimensions[:, 1])).all())

    def torus_correct(self, point) -> np.ndarray:
        return self.dimensions[:, 0] + np.mod(np.asanyarray(point) - self.dimensions[:, 0], self.size)

    # -- queries ------------------------------------------------------------------
    def _require_2d(self, what: str) -> None:
        if self.ndims != 2:
            raise NotImplementedError(f"{what} runs on the two-dimensional cell list; a {self.ndims}-dimensional space "
                                      "answers point queries (calculate_distances, get_agents_in_radius, ...) only")

    def _point_distances(self, point, sel, distances: bool, deltas: bool):
        p = np.asarray(point, dtype=np.float64).reshape(-1)
        if p.shape[0] != self.ndims:
            raise ValueError(f"point must have {self.ndims} components")
        if self.ndims == 2:
            return ops.point_distances(self.agent_positions.contiguous(), p, self.size, self.torus, sel=sel,
                                       distances=distances, deltas=deltas)
        return ops.point_distances_nd(self.agent_positions.contiguous(), p, self.size, self.torus, sel=sel,
                                      distances=distances, deltas=deltas)

    def _workspace(self, radius: float) -> ops.CellListWorkspace:
        if self._ws is None or self._ws.radius > radius or self._ws.n < self._n_agents:
            self._ws = ops.CellListWorkspace(max(self._n_agents, 1), self.size, radius, self.device)
        return self._ws

    def radius_query(self, points, radius: float, exclude: torch.Tensor | None = None):
        """Batched get_agents_in_radius for `points` [q, 2]: CSR (offsets, agent indices, float64 distances)."""
        self._require_2d("radius_query")
        q = torch.as_tensor(np.asarray(points, dtype=np.float32)) if not isinstance(points, torch.Tensor) else points
        q = q.to(self.device, torch.float32).contiguous()
        return ops.radius_query(self.agent_positions.contiguous(), q, self.size, float(radius), self.torus,
                                origin=self.dimensions[:, 0], exclude=exclude)

    def neighbors_in_radius(self, radius: float):
        """get_neighbors_in_radius of EVERY agent (self excluded) in one launch: CSR over agents."""
        n = self._n_agents
        me = torch.arange(n, dtype=torch.int32, device=self.device)
        return self.radius_query(self.agent_positions, radius, exclude=me)

    def get_agents_in_radius(self, point, radius: float | int = 1):
        """continuous_space.py:237-247: (agent indices, distances) with dist <= radius, storage order."""
        if self.ndims != 2:  # all n distances in one launch, then the reference's `distances <= radius` selection
            d, _ = self._point_distances(point, None, True, False)
            inside = ops.nonzero_indices(d <= float(radius))
            return inside.cpu().tolist(), d[inside].cpu().numpy()
        _, idx, dist = self.radius_query(np.asarray(point, dtype=np.float32)[None, :], radius)
        return idx.cpu().tolist(), dist.cpu().numpy()

    def _select(self, agents):
        if agents is None:
            return None, list(range(self._n_agents))
        ids = [int(a) for a in agents]
        return torch.as_tensor(ids, dtype=torch.int32, device=self.device), ids

    def calculate_distances(self, point, agents=None):
        """continuous_space.py:202-235 for one point against all (or the given) agents; float64."""
        sel, ids = self._select(agents)
        d, _ = self._point_distances(point, sel, True, False)
        return d.cpu().numpy(), ids

    def calculate_difference_vector(self, point, agents=None) -> np.ndarray:
        """continuous_space.py:169-200 (torus: the strictly smaller of delta and delta - sign(delta)*size)."""
        sel, _ = self._select(agents)
        _, v = self._point_distances(point, sel, False, True)
        return v.cpu().numpy()

    def k_nearest_query(self, points, k: int, exclude: torch.Tensor | None = None):
        """Batched get_k_nearest_agents for `points` [q, 2] in one launch: (indices [q, k], distances [q, k]) on
        the device, rows ascending in (distance, storage index), padded with -1 / inf."""
        self._require_2d("k_nearest_query")
        q = torch.as_tensor(np.asarray(points, dtype=np.float32)) if not isinstance(points, torch.Tensor) else points
        q = q.to(self.device, torch.float32).contiguous()
        return ops.knn_query(self.agent_positions.contiguous(), q, self.size, int(k), self.torus,
                             origin=self.dimensions[:, 0], exclude=exclude)

    def nearest_neighbors(self, k: int):
        """get_nearest_neighbors(k) of EVERY agent (self excluded, continuous_space_agents.py:80-91) in one launch."""
        me = torch.arange(self._n_agents, dtype=torch.int32, device=self.device)
        return self.k_nearest_query(self.agent_positions, k, exclude=me)

    def get_k_nearest_agents(self, point, k: int = 1):
        """continuous_space.py:249-283 (ties: earlier storage index first)."""
        n = self._n_agents
        if n == 0 or k <= 0:
            return [], np.array([])
        if k > n:
            warnings.warn(f"Requested k={k} nearest agents but only {n} agent(s) are present in the space. "
                          f"Returning all {n} agent(s).", UserWarning, stacklevel=2)
            k = n
        p = np.asarray(point, dtype=np.float64)
        if self.ndims != 2:  # the reference's own method: all n distances, keep the k smallest (ties: earlier index)
            d, _ = self._point_distances(p, None, True, False)
            order = ops.argsort_stable(d)[:k]
            return order.cpu().tolist(), d[order].cpu().numpy()
        if self.torus and not self.in_bounds(p):
            p = self.torus_correct(p)
        idx, dist = self.k_nearest_query(p[None, :], k)
        return idx[0].cpu().tolist(), dist[0].cpu().numpy()


ContinuousSpace = DeviceContinuousSpace


class DeviceContinuousSpaceAgent(DeviceAgent):
    """continuous_space_agents.py:19-91 on a row proxy: the agent's row is its storage index in `model.space` (the
    reference's `_mesa_index`).  Per-agent calls are debugging speed; the batch step uses
    `space.neighbors_in_radius` / `space.nearest_neighbors` / the Boids kernels."""

    __slots__ = ()

    @property
    def space(self) -> DeviceContinuousSpace:
        return self.model.space

    def __setattr__(self, name, value):
        if name == "position":  # continuous_space_agents.py:36-44: wrap on a torus, refuse out-of-bounds points otherwise
            space = self.space
            value = np.asarray(value, dtype=float)
            if not space.in_bounds(value):
                if space.torus:
                    value = space.torus_correct(value)
                else:
                    raise ValueError(f"point {value} is outside the bounds of the space")
        super().__setattr__(name, value)

    def _others(self, indices, dists):
        me = self._row()
        keep = np.asarray([i != me for i in indices], dtype=bool)
        agents = [self._set.agent_type(self._set, i) for i in indices if i != me]
        return agents, np.asarray(dists)[keep]

    def get_neighbors_in_radius(self, radius: float | int = 1):
        """continuous_space_agents.py:66-78: (agents, distances) within `radius` of this agent, itself dropped."""
        return self._others(*self.space.get_agents_in_radius(self.position, radius=radius))

    def get_nearest_neighbors(self, k: int = 1):
        """continuous_space_agents.py:80-91: the k nearest other agents (asks for k + 1 and drops itself)."""
        return self._others(*self.space.get_k_nearest_agents(self.position, k=k + 1))
