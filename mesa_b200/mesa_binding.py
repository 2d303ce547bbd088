"""The binding into a LIVE mesa installation (executed where `mesa` is importable: the build container; the GPU box
has no mesa, the self-contained mirror classes are what run there).

What a maintainer adds to use the device backend from an unmodified `mesa.Model`:

* `DeviceBackedModel(mesa.Model)` -- mesa/model.py:143-148 creates `_all_agents` / `_agents_by_type` as
  `_HardKeyAgentSet`s and `model.agents` / `model.agents_by_type` return them (model.py:192-247);
  `adopt(agentset)` swaps a `DeviceAgentSet` in, so that the model's own `self.agents.shuffle_do("step")`,
  `self.agents.do(...)`, `self.agents_by_type[T]`, `model.agents.get / agg / select` run on the device columns.
  The event-scheduled clock, `run_for` / `run_until`, scenarios and the data registry stay mesa's.
* `DiscreteSpace.register(DeviceGrid)` -- mesa/discrete_space/discrete_space.py:34-94: the device grids are
  (virtual) `DiscreteSpace`s, `AbstractAgentSet.register(DeviceAgentSet)` likewise (mesa/agentset.py:27-36:
  "Subclasses are free to override methods with optimized implementations").
"""
from __future__ import annotations

import mesa
import mesa.exceptions
from mesa.agentset import AbstractAgentSet
from mesa.discrete_space import DiscreteSpace

from .agentset import DeviceAgentSet
from . import space as _space
from .space import DeviceGrid, HostSpaceMirror

AbstractAgentSet.register(DeviceAgentSet)
DiscreteSpace.register(DeviceGrid)
DiscreteSpace.register(HostSpaceMirror)   # abstract_renderer.py:38-41 picks agent.cell.position for DiscreteSpace instances
# the device grids raise mesa's own exception type (mesa/exceptions.py) even if they were imported before mesa was
_space.CellMissingException = mesa.exceptions.CellMissingException


class DeviceBackedModel(mesa.Model):
    """A `mesa.Model` whose agent registry is a device population."""

    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        # the key of the device's Philox streams (mesa_b200/model.py): the stdlib seed mesa derived for model.random
        self.philox_seed = int(self.scenario._stdlib_seed) & 0xFFFFFFFFFFFFFFFF
        self._epoch = 0

    def _next_epoch(self) -> int:
        """Activation counter: every shuffle_do / shuffle is a fresh, replayable Philox permutation."""
        e = self._epoch
        self._epoch += 1
        return e

    def adopt(self, agentset: DeviceAgentSet, all_agents: bool = True) -> DeviceAgentSet:
        """Register a device population: `agents_by_type[agentset.agent_type]` and (all_agents=True) `model.agents`."""
        if not isinstance(agentset, DeviceAgentSet):
            raise TypeError("adopt() takes a DeviceAgentSet")
        agentset.model = self
        self._agents_by_type[agentset.agent_type] = agentset
        if all_agents:
            self._all_agents = agentset
        return agentset

    register_agentset = adopt

    def remove_all_agents(self):
        """model.py:307-320 for device populations (no per-agent Python object to call `remove` on)."""
        for agentset in list(self._agents_by_type.values()):
            if isinstance(agentset, DeviceAgentSet):
                agentset.clear()
            else:
                for agent in list(agentset):
                    agent.remove()
