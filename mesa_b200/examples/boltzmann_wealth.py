"""Boltzmann wealth model on the device (mirror of mesa/examples/basic/boltzmann_wealth_model/model.py:19-101,
agents.py:4-44): agents random-walk on a clamped Moore grid and give one unit to a random cell mate."""
from __future__ import annotations

import numpy as np
import torch

from .. import ops, reference_init
from ..agentset import DeviceAgent, DeviceAgentSet, batch_operator
from ..model import DeviceModel
from ..space import DeviceMooreGrid
from ._scenario import Scenario


class BoltzmannScenario(Scenario):
    """boltzmann_wealth_model/model.py:19-24."""

    n: int = 100
    width: int = 10
    height: int = 10


class MoneyAgent(DeviceAgent):
    """An agent with fixed initial wealth (agents.py:4)."""


@batch_operator(MoneyAgent, "step")
def _step(agents: DeviceAgentSet, epoch: int, **_):
    m = agents.model
    ops.boltzmann_step(agents.columns["cell"], agents.columns["wealth"], m._ws, m.philox_seed, epoch, torus=m.grid.torus)


class BoltzmannWealth(DeviceModel):
    """A simple model of an economy where agents exchange currency at random."""

    def __init__(self, scenario: BoltzmannScenario = BoltzmannScenario, initial_cells=None, device="cuda"):
        if isinstance(scenario, type):
            scenario = scenario()
        super().__init__(scenario=scenario)
        self.num_agents = scenario.n
        self.grid = DeviceMooreGrid((scenario.width, scenario.height), random=self.random, device=device)
        dev = self.grid.device
        ncells = scenario.width * scenario.height
        if initial_cells is None:
            # model.py:70-74 model.random.choices(cells, k=n): cells[floor(random() * ncells)] -- the same seed places
            # the agents where the reference does (mesa_b200/reference_init.py)
            if self.num_agents <= reference_init.MAX_REFERENCE_DRAWS:
                initial_cells = reference_init.boltzmann_initial_cells(self.random, self.num_agents, ncells)
            else:
                initial_cells = np.floor(self.rng.random(self.num_agents) * ncells).astype(np.int64)
        cell = torch.as_tensor(np.asarray(initial_cells)).to(torch.int32).to(dev).contiguous()
        wealth = torch.ones(self.num_agents, dtype=torch.int32, device=dev)
        self.agents = DeviceAgentSet(self, MoneyAgent, self.num_agents, {"cell": cell, "wealth": wealth}, self.random, resizable=False)
        self.grid.attach_population(self.agents)          # grid.agents / host mirror enumerate them through `cell`
        self._ws = ops.BoltzmannWorkspace(cell, scenario.width, scenario.height)
        self.running = True
        # grid.py:128 + cell.py:143-170: the implicit `empty` layer follows the occupancy.  Nobody on the step path reads
        # it, so it is brought up to date when somebody asks for it (hook), from the per-agent cell column
        self.grid.property_layers.hooks["empty"] = self._sync_empty

    def _sync_empty(self):
        empty = dict.__getitem__(self.grid.property_layers, "empty")
        empty.fill_(True)
        empty.view(-1)[self.agents.columns["cell"].long()] = False

    @property
    def last_iterations(self) -> int:
        """Rounds of the give-decision fixed point of the last step (device word, read on demand)."""
        self._ws.check()
        return int(self._ws.iterations.item())

    def step(self):
        self.agents.shuffle_do("step")  # model.py:79-81

    @property
    def gini(self) -> float:
        """model.py:83-101 on the device."""
        w = self.agents.columns["wealth"]
        x = w[ops.argsort_stable(w)].to(torch.float64)
        n = self.num_agents
        i = torch.arange(n, device=x.device, dtype=torch.float64)
        b = float((x * (n - i)).sum() / (n * x.sum()))
        return 1 + (1 / n) - 2 * b
