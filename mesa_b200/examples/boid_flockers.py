"""Boids flocking on the device (mirror of mesa/examples/basic/boid_flockers/model.py:22-120, agents.py:12-92).

ACTIVATION MODE (`BoidFlockers(activation=...)`): the reference's `shuffle_do("step")` is sequential -- each boid
sees the boids moved before it.  "sequential" reproduces exactly that in the Philox activation order
(ops.boids_step_sequential, csrc/continuous.cu) and is the DEFAULT, so the drop-in example follows the reference's
trajectory semantics; "synchronous" is the explicit fast mode (one pass): every boid reads the pre-step positions /
directions, i.e. the reference's own `Boid.step` arithmetic applied to a frozen snapshot."""
from __future__ import annotations

import numpy as np
import torch

from .. import ops
from ..agentset import DeviceAgent, DeviceAgentSet, batch_operator
from ..continuous_space import DeviceContinuousSpace, DeviceContinuousSpaceAgent
from ..model import DeviceModel
from ._scenario import Scenario


class BoidsScenario(Scenario):
    """boid_flockers/model.py:22-46."""

    population_size: int = 100
    width: int = 100
    height: int = 100
    speed: float = 1.0
    vision: float = 10.0
    separation: float = 2.0
    cohere: float = 0.03
    separate: float = 0.015
    match: float = 0.05


class Boid(DeviceContinuousSpaceAgent):
    """A Boid-style flocker agent (agents.py:12): a ContinuousSpaceAgent, so `boid.get_neighbors_in_radius(r)`,
    `boid.get_nearest_neighbors(k)` and `boid.position = p` (wrapped on the torus) work on a row proxy."""


@batch_operator(Boid, "step")
def _step(agents: DeviceAgentSet, epoch: int | None = None, **_):
    m = agents.model
    s = m.scenario
    pos, dirs = m.space._agent_positions, m._direction
    kw = dict(vision=s.vision, separation=s.separation, cohere=s.cohere, separate=s.separate, match=s.match,
              speed=s.speed, torus=m.space.torus, origin=m.space.dimensions[:, 0], pos_out=m._pos_next,
              dir_out=m._dir_next, neighbor_count=m._nbr_count)
    if m.activation == "sequential":
        if epoch is None:
            epoch = m._next_epoch()
        _, _, m.last_passes = ops.boids_step_sequential(pos[: m.space._n_agents], dirs, m.space.size, seed=m.philox_seed,
                                                        epoch=epoch, reach=m._reach, ws=m._seq_ws, **kw)
    else:
        if m.resort_every and m._steps_since_resort >= m.resort_every:
            m.resort()
            pos, dirs = m.space._agent_positions, m._direction
        m._steps_since_resort += 1
        ops.boids_step(pos[: m.space._n_agents], dirs, m.space.size, ws=m._ws, **kw)
    # ping-pong: the new state becomes the space's positions / the agents' direction column
    m.space._agent_positions, m._pos_next = m._pos_next, m.space._agent_positions
    m._direction, m._dir_next = m._dir_next, m._direction
    agents.columns["position"] = m.space._agent_positions
    agents.columns["direction"] = m._direction


class BoidFlockers(DeviceModel):
    """Flocker model class. Handles agent creation, placement and scheduling."""

    def __init__(self, scenario: BoidsScenario = BoidsScenario, initial_positions=None, initial_directions=None,
                 device="cuda", activation: str = "sequential", resort_every: int = 8):
        if isinstance(scenario, type):
            scenario = scenario()
        if activation not in ("synchronous", "sequential"):
            raise ValueError("activation must be 'synchronous' or 'sequential'")
        super().__init__(scenario=scenario)
        self.activation = activation
        self.last_passes = 0
        # synchronous activation: every `resort_every` steps the storage order is rewritten in cell-list order (ops.boids_resort),
        # which keeps the step's gathers / scatters coalesced; the boids' unique ids travel with the rows (0 = never)
        self.resort_every = int(resort_every) if activation == "synchronous" else 0
        self._steps_since_resort = self.resort_every       # the first step sorts the freshly created (random) order
        n = scenario.population_size
        self.space = DeviceContinuousSpace([[0, scenario.width], [0, scenario.height]], torus=True,
                                           random=self.random, n_agents=n, device=device)
        dev = self.space.device
        if initial_positions is None:
            initial_positions = self.rng.random(size=(n, 2)) * self.space.size  # model.py:79-81
        if initial_directions is None:
            initial_directions = self.rng.uniform(-1, 1, size=(n, 2))           # model.py:82
        self.space.add_agents(initial_positions)
        self._direction = torch.as_tensor(np.asarray(initial_directions, dtype=np.float32)).to(dev).contiguous()
        self._pos_next = torch.empty_like(self.space._agent_positions)
        self._dir_next = torch.empty_like(self._direction)
        self._nbr_count = torch.zeros(n, dtype=torch.int32, device=dev)
        self._ws = self._seq_ws = None
        if activation == "sequential":
            # a boid never moves farther than this in one step (constant over the run, ops.boids_reach)
            self._reach = ops.boids_reach(self._direction, scenario.speed)
            self._seq_ws = ops.BoidsSeqWorkspace(n, self.space.size, scenario.vision, self._reach, dev)
        else:
            self._ws = ops.CellListWorkspace(n, self.space.size, scenario.vision, dev)
        self.agents = DeviceAgentSet(self, Boid, n, {"position": self.space._agent_positions,
                                                     "direction": self._direction}, self.random, resizable=False)
        self.agent_angles = torch.zeros(n, device=dev)
        self.average_heading = None
        self.update_average_heading()

    def resort(self) -> None:
        """Rewrite the population in cell-list order.  Rows are renumbered (proxies taken before become invalid, like after a
        removal); `agents.unique_ids()` keeps telling who is who."""
        n = self.space._n_agents
        ids = self.agents.unique_ids().contiguous()
        pos, dirs, ids, _ = ops.boids_resort(self.space._agent_positions[:n].contiguous(), self._direction, self.space.size,
                                             vision=self.scenario.vision, ids=ids, torus=self.space.torus,
                                             origin=self.space.dimensions[:, 0], ws=self._ws)
        if self.space._agent_positions.shape[0] == n:
            self.space._agent_positions = pos
        else:
            self.space._agent_positions[:n].copy_(pos)
        self._direction = dirs
        self.agents.columns["position"], self.agents.columns["direction"] = self.space._agent_positions, self._direction
        self.agents._pop.ids = ids
        self.agents._pop.generation += 1
        self._steps_since_resort = 0

    def calculate_angles(self):
        d = self._direction
        self.agent_angles = torch.rad2deg(torch.atan2(d[:, 0], d[:, 1]))  # model.py:100-105

    def update_average_heading(self):
        if len(self.agents) == 0:
            self.average_heading = 0
            return
        mean = self._direction.to(torch.float64).mean(dim=0)
        self.average_heading = float(torch.atan2(mean[1], mean[0]))  # model.py:107-111

    def step(self):
        """model.py:113-120."""
        self.agents.shuffle_do("step")
        self.update_average_heading()
        self.calculate_angles()
