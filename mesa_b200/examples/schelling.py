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
    if m._small is not None:
        # small grid: ONE launch for the pending shuffle_do("step") (if any) + this do("assign_state"), nothing read back
        epoch, m._pending_epoch = m._pending_epoch, None
        layers = m.grid.property_layers      # (plain dict reads: an access through the hooks marks the layers as observed)
        m._small.step(dict.__getitem__(layers, "type"), dict.__getitem__(layers, "happy"), dict.__getitem__(layers, "agent_id"),
                      m.philox_seed, epoch or 0, agent_cell=dict.__getitem__(agents.columns, "cell"),
                      empty_layer=dict.__getitem__(layers, "empty"), relocate=epoch is not None, happy_current=m._happy_current)
        m._happy_current = True      # the happy layer now IS the evaluation of the type layer (until somebody writes a layer)
        m._last_small = m._last_small or epoch is not None
        return
    g = m.grid
    ops.schelling_happy(g.type, m.homophily, m.radius, g.torus, out=g.happy, count=m._happy_count, table=m._table)


@batch_operator(SchellingAgent, "step")
def _step(agents: DeviceAgentSet, epoch: int, **_):
    m = agents.model
    if m._small is not None:
        # deferred: the relocation is fused with the do("assign_state") that follows it in Schelling.step (model.py:87-93);
        # anybody who looks at the layers or the agents' cells in between gets it flushed first (_flush_pending)
        m._flush_pending()
        m._pending_epoch = epoch
        return
    m._relocate(epoch)


class _Columns(dict):
    """`cell` and `type` are stored per agent; `happy` is gathered from the happy layer on demand."""

    def __init__(self, model, **cols):
        super().__init__(**cols)
        self._model = model

    def __getitem__(self, name):
        self._model._flush_pending()
        if name == "happy":
            return self._model.grid.happy.reshape(-1)[dict.__getitem__(self, "cell").long()]
        return dict.__getitem__(self, name)

    def __contains__(self, name):
        return name == "happy" or dict.__contains__(self, name)


class Schelling(DeviceModel):
    """Model class for the Schelling segregation model."""

    def __init__(self, scenario: SchellingScenario = SchellingScenario, initial_types=None, device="cuda",
                 collect: bool = True, fused: bool = True):
        if isinstance(scenario, type):
            scenario = scenario()
        self._small = self._pending_epoch = None
        self._happy_current = False
        self._last_small = False
        self._running, self._running_pending = True, False
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
        self._last = (0, 0)
        if fused and ops.SchellingSmallStep.supports(int(types.numel())):
            # the whole step in one launch (csrc/schelling_small.cu); model.happy / movers stay on the device until read
            self._small = ops.SchellingSmallStep(types, n, self.homophily, self.radius, self.grid.torus)
            self._happy_count = self._small.out[0:1]
            for name in ("type", "happy", "agent_id", "empty"):
                self.grid.property_layers.hooks[name] = self._layer_observed
        cols = _Columns(self, cell=cells.to(torch.int32), type=flat[cells].clone())
        self.agents = DeviceAgentSet(self, SchellingAgent, n, cols, self.random, resizable=False)
        self.grid.cell_agents = self._cell_agents
        self.grid.attach_population(self.agents)
        self.agents.do("assign_state")  # model.py:83-85
        if self._small is None:
            self._sync_empty()          # grid.py:128 + cell.py:143-170: `empty` tracks the occupancy from placement on
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

    def _relocate(self, epoch: int) -> None:
        g = self.grid
        self._last_small = False
        self._last = ops.schelling_relocate(g.type, g.happy, g.agent_id, self._ws, self.philox_seed, epoch,
                                            agent_cell=dict.__getitem__(self.agents.columns, "cell"))

    def _layer_observed(self) -> None:
        """Somebody outside the step takes a layer tensor (and may write into it): a deferred relocation is run first, and
        the happy layer no longer counts as the evaluation of the type layer (a step without movers re-evaluates it)."""
        self._flush_pending()
        self._happy_current = False

    def _flush_pending(self) -> None:
        """A shuffle_do("step") that was deferred for fusion and is being observed before its do("assign_state"): run it
        now through the multi-launch path (same result)."""
        if self._pending_epoch is not None:
            epoch, self._pending_epoch = self._pending_epoch, None
            self._relocate(epoch)
            self._happy_current = False      # agents moved: the happy layer is stale until the next assign_state

    @property
    def happy(self) -> int:
        if self._small is not None:
            return self._small.check()[0]
        return int(self._happy_count.item())

    @property
    def last_movers(self) -> int:
        return self._small.check()[1] if self._last_small else self._last[0]

    @property
    def last_iterations(self) -> int:
        return self._small.check()[2] if self._last_small else self._last[1]

    # model.py:93 `self.running = self.happy < len(self.agents)`: evaluated when somebody reads it, so that a step
    # enqueues its launch and returns (run_for(n) never waits for the device; run_model() reads it every step)
    @property
    def running(self) -> bool:
        if self._running_pending:
            self._running, self._running_pending = self.happy < len(self.agents), False
        return self._running

    @running.setter
    def running(self, value) -> None:
        self._running, self._running_pending = bool(value), False

    def run_for(self, duration) -> None:
        """mesa/model.py:395-402.  On the single-launch path without a DataCollector the k steps that fall due run inside ONE
        launch (csrc/schelling_small.cu, nsteps = k, shuffle numbers of k consecutive epochs): a 100 x 100 step is shorter
        than the host's launch path, and nothing has to come back between the steps (`happy`, `running` are read lazily)."""
        import math

        until = self.time + duration
        k = int(math.floor(until) - math.floor(self.time))
        if self._small is None or self.datacollector is not None or k < 2 or type(self).step is not Schelling.step:
            return super().run_for(duration)
        self._flush_pending()
        epoch0 = self._epoch
        self._epoch += k
        layers = self.grid.property_layers
        self._small.step(dict.__getitem__(layers, "type"), dict.__getitem__(layers, "happy"), dict.__getitem__(layers, "agent_id"),
                         self.philox_seed, epoch0, agent_cell=dict.__getitem__(self.agents.columns, "cell"),
                         empty_layer=dict.__getitem__(layers, "empty"), relocate=True, nsteps=k, happy_current=True)
        self._last_small = True
        self.steps += k
        self.time = float(until)
        self._running_pending = True

    def step(self):
        """model.py:87-93."""
        self.agents.shuffle_do("step")
        self.agents.do("assign_state")
        if self._small is None:
            self._sync_empty()
        if self.datacollector is not None:
            self.datacollector.collect(self)  # model.py:92
        self._running_pending = True
