"""Schelling segregation on the device (mirror of mesa/examples/basic/schelling/model.py:9-93, agents.py:4-47).

State: int8 `type` layer (-1 empty cell, 0 majority, 1 minority), uint8 `happy` layer, int32 `agent_id`
layer (who sits where) and the per-agent `cell` column.  `do("assign_state")` is one stencil launch,
`shuffle_do("step")` the order-exact batched relocation (csrc/relocate.cu)."""
from __future__ import annotations

import numpy as np
import torch

from .. import ops, reference_init
from ..datacollection import DeviceDataCollector
from ..agentset import DeviceAgent, DeviceAgentSet, batch_operator
from ..model import DeviceModel
from ..space import DeviceMooreGrid
from ._scenario import Scenario


class SchellingScenario(Scenario):
    """schelling/model.py:9-28."""

    height: int = 20
    width: int = 20
    density: float = 0.8
    minority_pc: float = 0.5
    homophily: float = 0.4
    radius: int = 1


class SchellingAgent(DeviceAgent):
    """Schelling segregation agent (agents.py:4)."""

    @property
    def cell(self):
        k = int(self._set.columns["cell"][self._index].item())
        d1 = self.model.grid.dimensions[1]
        return self.model.grid[(k // d1, k % d1)]


@batch_operator(SchellingAgent, "assign_state")
def _assign_state(agents: DeviceAgentSet, **_):
    m = agents.model
    g = m.grid
    ops.schelling_happy(g.type, m.homophily, m.radius, g.torus, out=g.happy, count=m._happy_count, table=m._table)
    m._happy_dirty = True


@batch_operator(SchellingAgent, "step")
def _step(agents: DeviceAgentSet, epoch: int, **_):
    m = agents.model
    g = m.grid
    m.last_movers, m.last_iterations = ops.schelling_relocate(g.type, g.happy, g.agent_id, m._ws, m.philox_seed, epoch,
                                                              agent_cell=agents.columns["cell"])


class _Columns(dict):
    """`cell` and `type` are stored per agent; `happy` is gathered from the happy layer on demand."""

    def __init__(self, model, **cols):
        super().__init__(**cols)
        self._model = model

    def __getitem__(self, name):
        if name == "happy":
            return self._model.grid.happy.reshape(-1)[dict.__getitem__(self, "cell").long()]
        return dict.__getitem__(self, name)

    def __contains__(self, name):
        return name == "happy" or dict.__contains__(self, name)


class Schelling(DeviceModel):
    """Model class for the Schelling segregation model."""

    def __init__(self, scenario: SchellingScenario = SchellingScenario, initial_types=None, device="cuda",
                 collect: bool = True):
        if isinstance(scenario, type):
            scenario = scenario()
        super().__init__(scenario=scenario)
        # model.py:53-70: the reference's reporters, served from the device columns (one reduction / one bulk copy each)
        self.datacollector = DeviceDataCollector(
            model_reporters={
                "happy": "happy",
                "pct_happy": lambda m: (m.happy / len(m.agents)) * 100 if len(m.agents) > 0 else 0,
                "population": lambda m: len(m.agents),
                "minority_pct": lambda m: (int((m.agents.get("type") == 1).sum().item()) / len(m.agents) * 100
                                           if len(m.agents) > 0 else 0),
            },
            agent_reporters={"agent_type": "type"},
        ) if collect else None
        self.density, self.minority_pc = scenario.density, scenario.minority_pc
        self.homophily, self.radius = scenario.homophily, scenario.radius
        self.grid = DeviceMooreGrid((scenario.width, scenario.height), random=self.random, capacity=1, device=device)
        dev = self.grid.device
        if initial_types is None:
            # model.py:72-81 draws random() < density and then, for an agent, random() < minority_pc per cell, row-major
            # from model.random: the same seed gives the reference's initial grid (mesa_b200/reference_init.py)
            if 2 * scenario.width * scenario.height <= reference_init.MAX_REFERENCE_DRAWS:
                initial_types = reference_init.schelling_initial_types(self.random, scenario.width, scenario.height,
                                                                       self.density, self.minority_pc)
            else:  # sizes the reference cannot construct: the same distribution from the numpy stream
                occupied = self.rng.random((scenario.width, scenario.height)) < self.density
                minority = self.rng.random((scenario.width, scenario.height)) < self.minority_pc
                initial_types = np.where(occupied, minority.astype(np.int8), np.int8(-1))
        types = torch.as_tensor(np.asarray(initial_types, dtype=np.int8)).to(dev).contiguous()
        flat = types.reshape(-1)
        cells = torch.nonzero(flat >= 0).reshape(-1)  # registration order = row-major over occupied cells
        n = int(cells.numel())
        agent_id = torch.zeros(flat.numel(), dtype=torch.int32, device=dev)
        agent_id[cells] = torch.arange(n, dtype=torch.int32, device=dev)
        self.grid.add_property_layer("type", types)
        self.grid.create_property_layer("happy", 0, dtype=np.uint8)
        self.grid.add_property_layer("agent_id", agent_id.reshape(types.shape))
        self._table = ops.happy_threshold_table(self.homophily, self.radius)
        self._happy_count = torch.zeros(1, dtype=torch.int64, device=dev)
        self._ws = ops.RelocationWorkspace(types, n)
        self.last_movers = self.last_iterations = 0
        cols = _Columns(self, cell=cells.to(torch.int32), type=flat[cells].clone())
        self.agents = DeviceAgentSet(self, SchellingAgent, n, cols, self.random, resizable=False)
        self.grid.cell_agents = self._cell_agents
        self.agents.do("assign_state")  # model.py:83-85
        self._sync_empty()              # grid.py:128 + cell.py:143-170: `empty` tracks the occupancy from placement on
        if self.datacollector is not None:
            self.datacollector.collect(self)

    def _cell_agents(self, cells) -> list:
        """cell.py:192-195 on this capacity-1 grid: a cell's agent is the row its `agent_id` entry names."""
        flat = torch.as_tensor([c.flat for c in cells], dtype=torch.int64, device=self.grid.device)
        occupied = (self.grid.type.reshape(-1)[flat] >= 0).cpu().tolist()
        rows = self.grid.agent_id.reshape(-1)[flat].cpu().tolist()
        return [self.agents[r] for r, occ in zip(rows, occupied) if occ]

    # `grid.empty` follows the occupancy like the reference's implicit layer (grid.py:128, cell.py:143-170)
    def _sync_empty(self):
        torch.lt(self.grid.type, 0, out=self.grid.property_layers["empty"])

    @property
    def happy(self) -> int:
        return int(self._happy_count.item())

    def step(self):
        """model.py:87-93."""
        self.agents.shuffle_do("step")
        self.agents.do("assign_state")
        self._sync_empty()
        if self.datacollector is not None:
            self.datacollector.collect(self)  # model.py:92
        self.running = self.happy < len(self.agents)
