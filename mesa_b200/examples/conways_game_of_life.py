"""Conway's Game of Life on the device (mirror of mesa/examples/basic/conways_game_of_life/model.py:9-36,
agents.py:6-55).  One FixedAgent per cell on a torus Moore grid; `state` of agent i is cell i of the
uint8 `state` layer (row-major registration order, model.py:17-26)."""
from __future__ import annotations

import torch

from .. import ops
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
    ops.gol_step(m.grid.state, out=m._next_state, torus=m.grid.torus)


@batch_operator(Cell, "assume_state")
def _assume_state(agents: DeviceAgentSet, **_):
    # agents.py:53-55: state = _next_state -- a buffer swap, the layer tensor keeps its identity
    m = agents.model
    m._swap()


class ConwaysGameOfLife(DeviceModel):
    """Represents the 2-dimensional array of cells in Conway's Game of Life."""

    def __init__(self, width=50, height=50, initial_fraction_alive=0.2, rng=None, initial_state=None, device="cuda"):
        super().__init__(rng=rng)
        self.grid = DeviceMooreGrid((width, height), capacity=1, random=self.random, torus=True, device=device)
        if initial_state is None:
            initial_state = self.rng.random((width, height)) < initial_fraction_alive
        state = torch.as_tensor(initial_state).to(torch.uint8).to(self.grid.device).contiguous()
        self.grid.add_property_layer("state", state)
        self._next_state = torch.empty_like(state)
        self.agents = DeviceAgentSet(self, Cell, width * height, _StateColumns(self), self.random)
        self.running = True

    def _swap(self):
        cur = self.grid.property_layers["state"]
        self.grid.property_layers["state"] = self._next_state
        self._next_state = cur

    def step(self):
        """model.py:30-36: all cells compute their next state, then all cells assume it."""
        self.agents.do("determine_state")
        self.agents.do("assume_state")


class _StateColumns(dict):
    """columns view: `state` of agent i is element i of the current state layer."""

    def __init__(self, model):
        super().__init__()
        self._model = model

    def __getitem__(self, name):
        if name == "state":
            return self._model.grid.property_layers["state"].reshape(-1)
        raise KeyError(name)

    def __contains__(self, name):
        return name == "state"

    def values(self):
        return [self["state"]]
