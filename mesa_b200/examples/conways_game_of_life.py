"""Conway's Game of Life on the device (mirror of mesa/examples/basic/conways_game_of_life/model.py:9-36,
agents.py:6-55).  One FixedAgent per cell on a torus Moore grid; `state` of agent i is cell i of the
uint8 `state` layer (row-major registration order, model.py:17-26)."""
from __future__ import annotations

import torch

from .. import ops, reference_init
from ..agentset import DeviceAgent, DeviceAgentSet, batch_operator
from ..model import DeviceModel
from ..space import DeviceMooreGrid


class Cell(DeviceAgent):
    """Represents a single ALIVE or DEAD cell in the simulation (agents.py:6)."""

    DEAD = 0
    ALIVE = 1

    @property
    def is_alive(self):
        return self.state == self.ALIVE

    @property
    def x(self):
        return self._index // self.model.grid.dimensions[1]

    @property
    def y(self):
        return self._index % self.model.grid.dimensions[1]


@batch_operator(Cell, "determine_state")
def _determine_state(agents: DeviceAgentSet, **_):
    # agents.py:33-51 for every cell: next state into the second buffer (two-phase update)
    m = agents.model
    if m._bits is not None:
        m._sync_bits()
        ops.gol_step_bits(m._bits.cur, m._bits.nxt, 1, torus=m.grid.torus)
    else:
        ops.gol_step(m._layer(), out=m._next_state, torus=m.grid.torus)


@batch_operator(Cell, "assume_state")
def _assume_state(agents: DeviceAgentSet, **_):
    # agents.py:53-55: state = _next_state -- a buffer swap
    m = agents.model
    m._swap()


class ConwaysGameOfLife(DeviceModel):
    """Represents the 2-dimensional array of cells in Conway's Game of Life.

    `packed` (default: whenever the second dimension is a multiple of 32) keeps the state bit-packed on the
    device between reads of the `state` layer (csrc/gol_bits.cu): `step()` is one launch on 1/8 of the bytes and
    `run_for(n)` fuses several generations per launch.  Reading `grid.state` / `agents.get("state")` unpacks
    into the uint8 layer; writing into the layer tensor is picked up by the next step."""

    def __init__(self, width=50, height=50, initial_fraction_alive=0.2, rng=None, initial_state=None, device="cuda",
                 packed=None):
        super().__init__(rng=rng)
        self.grid = DeviceMooreGrid((width, height), capacity=1, random=self.random, torus=True, device=device)
        if initial_state is None:
            # model.py:19-27: row-major over the cells, alive iff model.random.random() < fraction -- the same seed gives
            # the reference's initial layer (mesa_b200/reference_init.py); beyond the sizes the reference can construct
            # the same distribution comes from the numpy stream
            if width * height <= reference_init.MAX_REFERENCE_DRAWS:
                initial_state = reference_init.gol_initial_state(self.random, width, height, initial_fraction_alive)
            else:
                initial_state = self.rng.random((width, height)) < initial_fraction_alive
        state = torch.as_tensor(initial_state).to(torch.uint8).to(self.grid.device).contiguous()
        self.grid.add_property_layer("state", state)
        if packed is None:
            packed = height % 32 == 0 and width >= 3
        self._bits = ops.BitLife(width, height, self.grid.device, torus=True) if packed else None
        self._bits_current = False   # the bit planes hold the current generation
        self._layer_current = True   # the uint8 layer holds the current generation
        self._next_state = None if packed else torch.empty_like(state)
        if packed:
            self.grid.property_layers.hooks["state"] = self._on_layer_access
        self.agents = DeviceAgentSet(self, Cell, width * height, _StateColumns(self), self.random, resizable=False)
        self.grid.cell_agents = self._cell_agents
        # one FixedAgent in every cell (model.py:13-26): no cell is ever empty (grid.py:128, cell.py:143-170)
        dict.__getitem__(self.grid.property_layers, "empty").fill_(False)
        self.running = True

    def _cell_agents(self, cells) -> list:
        """One FixedAgent per cell, created row-major (model.py:13-26): the agent's row is the flat cell index."""
        return [self.agents[c.flat] for c in cells]

    # -- the two representations of the state layer
    def _layer(self, writable: bool = True) -> torch.Tensor:
        layer = dict.__getitem__(self.grid.property_layers, "state")
        if self._bits is not None:
            if not self._layer_current:
                self._bits.store(layer)
                self._layer_current = True
            if writable:  # the caller may edit the tensor in place: the bit planes are re-packed before the next step
                self._bits_current = False
        return layer

    def _on_layer_access(self):
        self._layer(writable=True)

    def _sync_bits(self):
        if not self._bits_current:
            self._bits.load(dict.__getitem__(self.grid.property_layers, "state"))
            self._bits_current = True

    def _swap(self):
        if self._bits is not None:
            b = self._bits
            b.cur, b.nxt = b.nxt, b.cur
            self._layer_current = False
            return
        cur = dict.__getitem__(self.grid.property_layers, "state")
        dict.__setitem__(self.grid.property_layers, "state", self._next_state)
        self._next_state = cur

    def step(self):
        """model.py:30-36: all cells compute their next state, then all cells assume it."""
        self.agents.do("determine_state")
        self.agents.do("assume_state")

    def run_for(self, duration: int) -> None:
        """`duration` steps; with the packed state several generations share one launch (identical result)."""
        if self._bits is None:
            return super().run_for(duration)
        n = int(duration)
        if n <= 0:
            return
        self._sync_bits()
        self._bits.run(n)
        self._layer_current = False
        self.steps += n
        self.time += float(n)


class _StateColumns(dict):
    """columns view: `state` of agent i is element i of the current state layer."""

    def __init__(self, model):
        super().__init__()
        self._model = model

    def __getitem__(self, name):
        if name == "state":
            return self._model._layer(writable=False).reshape(-1)
        raise KeyError(name)

    def __contains__(self, name):
        return name == "state"

    def values(self):
        return [self["state"]]
