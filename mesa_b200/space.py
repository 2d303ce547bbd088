"""Device-resident orthogonal grids: the drop-in mirror of mesa.discrete_space.Grid for the hot path.

Mirrors the public surface of
    mesa/discrete_space/grid.py:59-128   Grid(dimensions, torus, capacity, random, cell_klass)
    grid.py:434-501                      OrthogonalMooreGrid / OrthogonalVonNeumannGrid
    grid.py:130-234                      create/add/remove_property_layer, get_neighborhood_mask
    grid.py:288-316                      select_random_empty_cell
    discrete_space.py:89-94, 169-183     find_nearest_cell, all_cells, __getitem__
    cell.py:37, 192-276                  Cell.coordinate / is_empty / neighborhood / get_neighborhood
with the same argument meaning and error behaviour (ValueError / TypeError / CellMissingException),
but WITHOUT one Python object per cell: state is a set of [d0, d1] device layers (torch tensors),
cells are lightweight proxies created on demand, and whole-grid neighbourhood work is one
kernel launch (ops.neighborhood_sum, ops.schelling_happy, ops.gol_step).

Grids of 1 to 4 dimensions (grid.py:457-462, 488-501 for n != 2); the tuned stencil kernels and the row-sharded
multi-GPU paths are two-dimensional (the BASELINE.json configs), other dimensionalities use the generic
neighbourhood kernel (mb_neighborhood_sum_nd).
"""
from __future__ import annotations

from itertools import product
from random import Random

import numpy as np
import torch

from . import ops

try:  # the reference's exception types when mesa is importable, same names otherwise
    from mesa.exceptions import CellMissingException  # type: ignore
except Exception:  # pragma: no cover - mesa is not installed on the GPU box
    class CellMissingException(KeyError):
        """Raised when a coordinate is not part of the grid (mesa/exceptions.py)."""

        def __init__(self, coordinate):
            super().__init__(f"Cell at coordinate {coordinate} does not exist.")
            self.coordinate = coordinate


_TORCH_DTYPES = {
    float: torch.float64, int: torch.int32, bool: torch.bool,
    np.float64: torch.float64, np.float32: torch.float32, np.int32: torch.int32, np.int64: torch.int64,
    np.uint8: torch.uint8, np.int8: torch.int8, np.bool_: torch.bool,
}


def _torch_dtype(dtype):
    if isinstance(dtype, torch.dtype):
        return dtype
    if dtype in _TORCH_DTYPES:
        return _TORCH_DTYPES[dtype]
    return _TORCH_DTYPES[np.dtype(dtype).type]


class _LayerDict(dict):
    """`Grid.property_layers` (grid.py:112): a plain dict of [d0, d1] tensors, plus optional per-layer access hooks
    so that a model may keep a layer in another device format between reads (e.g. the bit-packed Game of Life
    state) and bring the tensor up to date only when somebody asks for it."""

    def __init__(self):
        super().__init__()
        self.hooks: dict = {}

    def __getitem__(self, name):
        hook = self.hooks.get(name)
        if hook is not None:
            hook()
        return super().__getitem__(name)

    def get(self, name, default=None):
        return self[name] if name in self else default

    def values(self):
        return [self[k] for k in self]

    def items(self):
        return [(k, self[k]) for k in self]


class CellProxy:
    """A cell of a DeviceGrid: coordinate + accessors that read the device layers (debug speed)."""

    __slots__ = ("grid", "coordinate")

    def __init__(self, grid: "DeviceGrid", coordinate: tuple[int, ...]):
        object.__setattr__(self, "grid", grid)
        object.__setattr__(self, "coordinate", coordinate)

    # property layers as attributes (grid.py:206-220)
    def __getattr__(self, name):
        layers = object.__getattribute__(self, "grid").property_layers
        if name in layers:
            return layers[name][self.coordinate].item()
        raise AttributeError(name)

    def __setattr__(self, name, value):
        grid = self.grid
        if name in grid.property_layers:
            if name in grid._read_only_layers:
                raise AttributeError(f"property_layer '{name}' is read-only")
            grid.property_layers[name][self.coordinate] = value
            return
        raise AttributeError(name)

    def __eq__(self, other):
        return isinstance(other, CellProxy) and other.grid is self.grid and other.coordinate == self.coordinate

    def __hash__(self):
        return hash(self.coordinate)

    def __repr__(self):
        return f"Cell({self.coordinate})"

    @property
    def flat(self) -> int:
        return int(np.ravel_multi_index(self.coordinate, self.grid.dimensions))

    @property
    def position(self) -> np.ndarray:
        """cell.py:68-78: the physical position of a cell of an implicit grid is its coordinate."""
        return np.asarray(self.coordinate, dtype=float)

    @property
    def is_empty(self) -> bool:
        return bool(self.grid.empty[self.coordinate].item())

    @property
    def is_full(self) -> bool:
        """cell.py:183-189: len(agents) >= capacity (never full when capacity is None)."""
        grid = self.grid
        if grid.capacity is None:
            return False
        return bool((grid.agent_counts()[self.coordinate] >= grid.capacity).item())

    @property
    def connections(self) -> dict:
        """cell.py:96: {offset: neighbouring cell} in the order the grid connects them (grid.py:390-399; on a torus an offset
        that wraps onto an already connected cell keeps its own key, like the reference's dict of connections)."""
        grid, out = self.grid, {}
        for off in grid._offsets:
            nb = tuple(x + o for x, o in zip(self.coordinate, off))
            if grid.torus:
                nb = tuple(x % d for x, d in zip(nb, grid.dimensions))
            if all(0 <= x < d for x, d in zip(nb, grid.dimensions)):
                out[off] = CellProxy(grid, nb)
        return out

    @property
    def neighborhood(self) -> list["CellProxy"]:
        return self.get_neighborhood()

    def get_neighborhood(self, radius: int = 1, include_center: bool = False) -> list["CellProxy"]:
        """Neighbour cells in the reference's order (connection order for radius 1, BFS beyond) as a
        cell collection (cell.py:202-223 returns a CellCollection)."""
        cells = [CellProxy(self.grid, c) for c in self.grid.neighborhood_coordinates(self.coordinate, radius, include_center)]
        return CellCollectionProxy(cells, getattr(self.grid, "random", None))

    @property
    def agents(self) -> list:
        """cell.py:192-195: the agents in this cell, through the model's `grid.cell_agents` hook."""
        return CellCollectionProxy([self], None).agents


_RAISES = object()
# cell.py:49-58 `Cell.__slots__` (+ the two fields of the proxy): names a property layer may not take
_CELL_SLOTS = frozenset({"_agents", "_empty", "_position", "capacity", "connections", "coordinate", "properties", "random",
                         "grid"})


class CellCollectionProxy(list):
    """An ordered collection of cell proxies with the reference CellCollection's methods
    (cell_collection.py:33-175): `cells`, `agents`, `select_random_cell`, `select_random_agent`, `select`.  It IS
    the list of its cells, so code that indexes or compares the result of `get_neighborhood` keeps working.
    The random draws are the reference's (`random.choice` on the ordered cells / agents with the grid's generator),
    so a seeded host generator gives the reference's picks."""

    def __init__(self, cells=(), random: Random | None = None):
        super().__init__(cells)
        self.random = random

    @property
    def cells(self) -> list:
        return list(self)

    @property
    def agents(self) -> list:
        """Agents of the cells, cell by cell in collection order (cell_collection.py:97-99).  Which rows of the
        model's population sit in a cell is model state (an `agent_id` layer, a per-agent `cell` column ...), so the
        model provides `grid.cell_agents(list of cells) -> list of agents`."""
        if not self:
            return []
        hook = getattr(self[0].grid, "cell_agents", None)
        if hook is None:
            raise NotImplementedError("this grid's model did not register grid.cell_agents (cells -> agents)")
        return list(hook(list(self)))

    def select_random_cell(self, default=_RAISES):
        if not self:
            if default is _RAISES:
                raise LookupError("Cannot select random cell from empty collection")
            return default
        return self.random.choice(self.cells)

    def select_random_agent(self, default=_RAISES):
        agents = self.agents
        if not agents:
            if default is _RAISES:
                raise LookupError("Cannot select random agent from empty collection")
            return default
        return self.random.choice(agents)

    def select(self, filter_func=None, at_most=float("inf")) -> "CellCollectionProxy":
        """cell_collection.py:138-175: the first `at_most` cells passing `filter_func` (a float <= 1.0 is a fraction of
        the collection, rounded down)."""
        if filter_func is None and at_most == float("inf"):
            return self
        if at_most <= 1.0 and isinstance(at_most, float):
            at_most = int(len(self) * at_most)
        out = []
        for cell in self:
            if len(out) >= at_most:
                break
            if not filter_func or filter_func(cell):
                out.append(cell)
        return CellCollectionProxy(out, self.random)

    def __repr__(self):
        return f"CellCollection({list(self)})"


class HostAgentView:
    """One agent of a host mirror: `unique_id`, `cell` (a cell proxy: `.coordinate`, `.position`) and the mirrored
    attribute values -- what a portrayal function or a checkpoint inspector reads (visualization/backends/
    matplotlib_backend.py:104-163: `for agent in space.agents: agent_portrayal(agent)`, `agent.cell.position`)."""

    __slots__ = ("unique_id", "cell", "agent_type", "_attrs")

    def __init__(self, unique_id, cell, agent_type, attrs):
        self.unique_id, self.cell, self.agent_type, self._attrs = unique_id, cell, agent_type, attrs

    def __getattr__(self, name):
        try:
            return object.__getattribute__(self, "_attrs")[name]
        except KeyError:
            raise AttributeError(name) from None

    def __repr__(self):
        return f"<{self.agent_type.__name__} {self.unique_id} at {self.cell.coordinate}>"


class HostSpaceMirror:
    """A read-only HOST snapshot of a device grid for consumers that walk Python objects (SURVEY.md 8(f) next-3:
    visualization/space_renderer.py:339, backends/matplotlib_backend.py:104, 340; checkpoint inspection): every property
    layer as a numpy array (`space.property_layers`), every agent of the populations attached to the grid as a
    HostAgentView (`space.agents`), plus the geometry the renderers read.  Built with ONE device-to-host copy per layer and
    per agent column; the device state is not touched afterwards."""

    def __init__(self, grid: "DeviceGrid", layers=None, attributes=None):
        self.dimensions, self.torus, self.capacity = grid.dimensions, grid.torus, grid.capacity
        self.width = grid.dimensions[0]
        self.height = grid.dimensions[1] if len(grid.dimensions) > 1 else 1
        names = list(grid.property_layers) if layers is None else list(layers)
        self.property_layers = {}
        for name in names:
            t = grid.property_layers[name]
            self.property_layers[name] = t.detach().to("cpu", copy=True).numpy()
        self._grid = grid
        self.agents = []
        for agentset, column in grid.populations:
            cols = agentset.columns
            cells = cols[column].detach().cpu().numpy().astype(np.int64).reshape(-1)[: agentset.n]
            ids = agentset.unique_ids_host()
            wanted = [k for k in cols if k != column] if attributes is None else [k for k in attributes if k in cols]
            host = {k: cols[k].detach().cpu().numpy()[: agentset.n] for k in wanted}
            coords = np.stack(np.unravel_index(cells, grid.dimensions), axis=1) if len(cells) else np.zeros((0, len(grid.dimensions)), int)
            for r in range(agentset.n):
                attrs = {k: (v[r].item() if v[r].ndim == 0 else v[r]) for k, v in host.items()}
                self.agents.append(HostAgentView(int(ids[r]), CellProxy(grid, tuple(int(c) for c in coords[r])),
                                                 agentset.agent_type, attrs))

    @property
    def all_cells(self):
        return self._grid.all_cells

    def __getitem__(self, coordinate):
        return self._grid[coordinate]


class _CellMap:
    """`space._cells` (discrete_space.py:67): coordinate -> cell, without materialising the cells."""

    def __init__(self, grid: "DeviceGrid"):
        self._grid = grid

    def __getitem__(self, coordinate) -> CellProxy:
        return self._grid[coordinate]

    def __contains__(self, coordinate) -> bool:
        try:
            self._grid._check(coordinate)
        except KeyError:
            return False
        return True

    def __len__(self) -> int:
        return len(self._grid)

    def __iter__(self):
        return product(*(range(d) for d in self._grid.dimensions))

    def keys(self):
        return iter(self)

    def values(self):
        return self._grid.all_cells

    def items(self):
        return ((c, CellProxy(self._grid, c)) for c in self)


class _DetachedCell(CellProxy):
    """`grid.cell_klass(coordinate)` (grid.py:100-110): a cell that belongs to no grid yet.  Like the reference's slotted
    Cell it rejects unknown attributes."""

    __slots__ = ()

    def __init__(self, coordinate, capacity=None, random=None):
        CellProxy.__init__(self, None, coordinate if isinstance(coordinate, tuple) else (coordinate,))

    def __setattr__(self, name, value):
        raise AttributeError(f"'Cell' object has no attribute {name!r}")


class DeviceGrid:
    """Base class of the device grids (mirror of mesa.discrete_space.Grid)."""

    cell_klass = _DetachedCell

    @property
    def _cells(self) -> _CellMap:
        return _CellMap(self)

    _offsets: list[tuple[int, int]] = []
    _moore = True

    @property
    def width(self) -> int:
        return self.dimensions[0]

    @property
    def height(self) -> int:
        return self.dimensions[1]

    @staticmethod
    def _offsets_nd(ndims: int) -> list[tuple[int, ...]]:  # pragma: no cover - overridden
        raise NotImplementedError

    def __init__(self, dimensions, torus: bool = False, capacity: float | None = None,
                 random: Random | None = None, device="cuda"):
        self.torus = torus
        self.dimensions = tuple(dimensions) if not isinstance(dimensions, tuple) else dimensions
        self.capacity = capacity
        self._validate_parameters()
        if not 1 <= len(self.dimensions) <= 4:
            raise NotImplementedError("mesa_b200 device grids have 1 to 4 dimensions")
        self._ndims = len(self.dimensions)
        if self._ndims != 2:
            self._offsets = self._offsets_nd(self._ndims)
        self.random = random if random is not None else Random()
        self.device = torch.device(device)
        self.property_layers: dict[str, torch.Tensor] = _LayerDict()
        self._read_only_layers: set[str] = set()
        self.populations: list = []     # (DeviceAgentSet, name of its flat-cell column): who lives on this grid
        self.create_property_layer("empty", default_value=True, dtype=bool)

    # grid.py:280-286
    def _validate_parameters(self):
        if not all(isinstance(dim, int) and dim > 0 for dim in self.dimensions):
            raise ValueError("Dimensions must be a list of positive integers.")
        if not isinstance(self.torus, bool):
            raise TypeError("Torus must be a boolean.")
        if self.capacity is not None and not isinstance(self.capacity, float | int):
            raise TypeError("Capacity must be a number or None.")

    # ------------------------------------------------------------------ agents on the grid (readback)
    def attach_population(self, agentset, cell_column: str = "cell") -> None:
        """Tell the grid that `agentset` lives on it, its agents' flat cell ids in column `cell_column` (the device
        models do this; it is what `grid.agents` and the host mirror enumerate)."""
        if all(a is not agentset for a, _ in self.populations):
            self.populations.append((agentset, cell_column))

    def host_mirror(self, layers=None, attributes=None) -> "HostSpaceMirror":
        """A host snapshot for visualization / inspection: numpy layers + lightweight agent views (see HostSpaceMirror)."""
        return HostSpaceMirror(self, layers, attributes)

    @property
    def agents(self) -> list:
        """discrete_space.py:84-87 `space.agents`: every agent in the space, cell by cell in row-major cell order and,
        inside a cell, in registration order.  A host-side list (one bulk copy per population): debugging / rendering
        speed, not a step path."""
        views = self.host_mirror(layers=()).agents
        flat = [int(np.ravel_multi_index(a.cell.coordinate, self.dimensions)) for a in views]
        return [views[i] for i in sorted(range(len(views)), key=lambda i: flat[i])]

    # ------------------------------------------------------------------ property layers
    def create_property_layer(self, name: str, default_value=0.0, dtype=float, read_only: bool = False) -> torch.Tensor:
        layer = torch.empty(self.dimensions, dtype=_torch_dtype(dtype), device=self.device)
        if layer.dtype in (torch.bool, torch.uint8, torch.int32, torch.float32, torch.float64):
            ops.layer_fill(layer.view(torch.uint8) if layer.dtype == torch.bool else layer, default_value)
        else:
            layer.fill_(default_value)
        self._attach_property_layer(name, layer, read_only=read_only)
        return layer

    def add_property_layer(self, name: str, array, read_only: bool = False) -> None:
        if tuple(array.shape) != tuple(self.dimensions):
            raise ValueError(f"Array shape {tuple(array.shape)} does not match grid dimensions {self.dimensions}.")
        if not isinstance(array, torch.Tensor):
            array = torch.as_tensor(np.asarray(array)).to(self.device)
        elif array.device != self.device:
            array = array.to(self.device)
        self._attach_property_layer(name, array, read_only=read_only)

    def remove_property_layer(self, name: str) -> None:
        if name not in self.property_layers:
            raise KeyError(f"No property_layer named '{name}'.")
        del self.property_layers[name]
        self._read_only_layers.discard(name)

    def _attach_property_layer(self, name: str, layer: torch.Tensor, read_only: bool = False) -> None:
        if any(name in c.__dict__ for c in type(self).__mro__):
            raise ValueError(f"property_layer '{name}' conflicts with an existing Grid attribute.")
        if name in _CELL_SLOTS:   # grid.py:194-202: the name would shadow a field of the cells
            raise ValueError(f"property_layer name '{name}' clashes with existing slot '{name}'.")
        if name in self.property_layers:
            raise ValueError(f"property_layer '{name}' already exists.")
        self.property_layers[name] = layer
        if read_only:
            self._read_only_layers.add(name)

    def __getattr__(self, name):
        # grid.<layer> access (grid.py:205)
        layers = self.__dict__.get("property_layers")
        if layers is not None and name in layers:
            return layers[name]
        raise AttributeError(f"{type(self).__name__!s} has no attribute {name!r}")

    # ------------------------------------------------------------------ cells
    def _check(self, coordinate) -> tuple[int, ...]:
        try:
            coord = tuple(coordinate)
        except Exception:
            raise CellMissingException(coordinate) from None
        if len(coord) != len(self.dimensions) or not all(
                isinstance(c, (int, np.integer)) and 0 <= c < d for c, d in zip(coord, self.dimensions)):
            raise CellMissingException(coordinate)
        return tuple(int(c) for c in coord)

    def __getitem__(self, coordinate) -> CellProxy:
        return CellProxy(self, self._check(coordinate))

    @property
    def all_cells(self):
        """Row-major iterator over cell proxies (grid.py:120-126 order); lazy -- do not materialise large grids."""
        return (CellProxy(self, c) for c in product(*(range(d) for d in self.dimensions)))

    def __len__(self) -> int:
        return int(np.prod(self.dimensions))

    def _unravel(self, k: int) -> tuple[int, ...]:
        return tuple(int(c) for c in np.unravel_index(int(k), self.dimensions))

    @property
    def empties(self):
        idx = ops.nonzero_indices(self.empty).cpu().tolist()
        return [CellProxy(self, self._unravel(k)) for k in idx]

    def find_nearest_cell(self, position) -> CellProxy:
        """Floor of the position, wrapped on a torus (discrete_space tests test_spatial_lookups.py:20-44)."""
        pos = np.asarray(position, dtype=float)
        idx = np.floor(pos).astype(int)
        if self.torus:
            idx = idx % np.array(self.dimensions)
        if np.any(idx < 0) or np.any(idx >= np.array(self.dimensions)):
            raise ValueError(f"Position {position} is outside the grid bounds.")
        return CellProxy(self, tuple(int(c) for c in idx))

    # ------------------------------------------------------------------ neighbourhoods
    def neighborhood_coordinates(self, coordinate, radius: int = 1, include_center: bool = False) -> list[tuple[int, int]]:
        """Distinct neighbour coordinates of one cell (Cell._neighborhood, cell.py:225-276): BFS over
        the connection offsets, so Moore = Chebyshev, von Neumann = Manhattan; torus wraps and dedups."""
        if radius < 1:
            raise ValueError("radius must be at least 1")
        start = self._check(coordinate)
        dims = self.dimensions

        def connections(c):
            out = []
            for off in self._offsets:
                nb = tuple(x + o for x, o in zip(c, off))
                if self.torus:
                    nb = tuple(x % d for x, d in zip(nb, dims))
                if all(0 <= x < d for x, d in zip(nb, dims)):
                    out.append(nb)
            return out

        if radius == 1:
            seen = dict.fromkeys(connections(start))
            if include_center:
                seen[start] = None
            return list(seen)
        visited = {start}
        layer = connections(start)
        found: dict = {}
        for c in layer:
            if c not in visited:
                visited.add(c)
                found[c] = None
        for _ in range(radius - 1):
            nxt = []
            for c in layer:
                for nb in connections(c):
                    if nb not in visited:
                        visited.add(nb)
                        nxt.append(nb)
                        found[nb] = None
            layer = nxt
            if not layer:
                break
        if include_center:
            found[start] = None
        else:
            found.pop(start, None)
        return list(found)

    def get_neighborhood_mask(self, coordinate, include_center: bool = True, radius: int = 1) -> torch.Tensor:
        mask = torch.zeros(self.dimensions, dtype=torch.bool, device=self.device)
        coords = self.neighborhood_coordinates(coordinate, radius, include_center)
        if coords:
            idx = torch.tensor(coords, device=self.device)
            mask[tuple(idx[:, d] for d in range(idx.shape[1]))] = True
        return mask

    def neighborhood_sum(self, layer, radius: int = 1, include_center: bool = False) -> torch.Tensor:
        """Sum of a property layer over every cell's neighbourhood in one launch."""
        if isinstance(layer, str):
            layer = self.property_layers[layer]
        if layer.dtype == torch.bool:
            layer = layer.view(torch.uint8)
        return ops.neighborhood_sum(layer, radius, self._moore, include_center, self.torus)

    # ------------------------------------------------------------------ empties
    def select_random_empty_cell(self) -> CellProxy:
        """grid.py:288-316: <= 50 uniform probes, first empty wins; else argwhere(empty) + one choice."""
        ncells = len(self)
        empty = self.empty
        for _ in range(50):
            k = self.random._randbelow(ncells)
            if bool(empty.reshape(-1)[k].item()):
                return CellProxy(self, self._unravel(k))
        empties = ops.nonzero_indices(empty)
        if empties.numel() == 0:
            raise ValueError("Grid is completely full. No empty cells available. Cannot select a random empty cell.")
        k = int(empties[self.random._randbelow(empties.numel())].item())
        return CellProxy(self, self._unravel(k))


    # ------------------------------------------------------------------ capacity (grid.py:318-378)
    def agent_counts(self) -> torch.Tensor:
        """Agents per cell: the `agent_count` property layer when the model maintains one (cells shared by several
        agents), otherwise 1 - empty (grids of capacity 1, where `empty` is the occupancy)."""
        layers = self.property_layers
        if "agent_count" in layers:
            return layers["agent_count"]
        return (~layers["empty"]).to(torch.int32)

    @property
    def cells_with_capacity(self):
        """All cells that are not full (grid.py:318-350); every cell when capacity is None."""
        if self.capacity is None:
            return self.all_cells
        idx = ops.nonzero_indices(self.agent_counts() < self.capacity).cpu().tolist()
        return [CellProxy(self, self._unravel(k)) for k in idx]

    def select_random_cell_with_capacity(self) -> CellProxy:
        """grid.py:352-378: <= 50 uniform probes, first cell that is not full wins; else one choice among the
        cells with capacity; IndexError when every cell is full."""
        ncells = len(self)
        if self.capacity is None:
            return CellProxy(self, self._unravel(self.random._randbelow(ncells)))
        room = (self.agent_counts() < self.capacity).reshape(-1)
        for _ in range(50):
            k = self.random._randbelow(ncells)
            if bool(room[k].item()):
                return CellProxy(self, self._unravel(k))
        available = ops.nonzero_indices(room)
        if available.numel() == 0:
            raise IndexError("No available cells exist in the grid: all cells are at full capacity.")
        k = int(available[self.random._randbelow(available.numel())].item())
        return CellProxy(self, self._unravel(k))


class DeviceMooreGrid(DeviceGrid):
    """8-connected grid (OrthogonalMooreGrid, grid.py:434-462)."""

    # fmt: off
    _offsets = [(-1, -1), (-1, 0), (-1, 1),
                ( 0, -1),          ( 0, 1),
                ( 1, -1), ( 1, 0), ( 1, 1)]
    # fmt: on
    _moore = True

    @staticmethod
    def _offsets_nd(ndims: int) -> list[tuple[int, ...]]:
        """grid.py:457-462: product([-1, 0, 1]^n) without the centre."""
        offsets = list(product([-1, 0, 1], repeat=ndims))
        offsets.remove((0,) * ndims)
        return offsets


class DeviceVonNeumannGrid(DeviceGrid):
    """4-connected grid (OrthogonalVonNeumannGrid, grid.py:465-501)."""

    _offsets = [(-1, 0), (0, -1), (0, 1), (1, 0)]
    _moore = False

    @staticmethod
    def _offsets_nd(ndims: int) -> list[tuple[int, ...]]:
        """grid.py:488-501: one step back, one step forward along every axis."""
        offsets = []
        for dim in range(ndims):
            for delta in (-1, 1):
                off = [0] * ndims
                off[dim] = delta
                offsets.append(tuple(off))
        return offsets


# names a Mesa user would look for
OrthogonalMooreGrid = DeviceMooreGrid
OrthogonalVonNeumannGrid = DeviceVonNeumannGrid
