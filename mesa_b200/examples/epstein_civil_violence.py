"""Epstein's civil-violence model on the device (mirror of mesa/examples/advanced/epstein_civil_violence/model.py:16-125,
agents.py:7-140): citizens rebel when their grievance outweighs the risk of arrest they perceive in their radius-`vision`
neighbourhood, cops arrest active citizens in theirs, everybody moves to a random EMPTY cell of the neighbourhood.

The reference model's own step runs unchanged --

    self.agents.shuffle_do("step")          # citizens and cops in one shuffled order
    self._update_counts(); self.datacollector.collect(self)

-- on ONE device population (citizens and cops share the activation, `kind` tells them apart; `agents_by_type[Citizen]` /
`[Cop]` are views of it).  The sequential semantics are computed exactly by one launch of a rank-ordered dependency graph
(csrc/epstein.cu); the same seed gives the reference's initial state, and with the replayed Philox order / draws the
reference's trajectory (tests/golden/epstein_*.npz: cells, states, jail sentences and float64 arrest probabilities bit for
bit).
"""
from __future__ import annotations

import ctypes
import math
from enum import Enum

import numpy as np
import torch

from .. import _native as N
from ..agentset import DeviceAgent, DeviceAgentSet, batch_operator
from ..datacollection import DeviceDataCollector
from ..model import DeviceModel
from ..space import DeviceMooreGrid, DeviceVonNeumannGrid
from ._scenario import Scenario


class CitizenState(Enum):
    """agents.py:7-10."""

    ACTIVE = 1
    QUIET = 2
    ARRESTED = 3


class EpsteinScenario(Scenario):
    """model.py:16-31."""

    citizen_density: float = 0.7
    cop_density: float = 0.074
    citizen_vision: int = 7
    cop_vision: int = 7
    legitimacy: float = 0.8
    max_jail_term: int = 1000
    active_threshold: float = 0.1
    arrest_prob_constant: float = 2.3
    movement: bool = True
    max_iters: int = 1000
    activation_order: str = "Random"
    grid_type: str = "Von Neumann"
    rng = 42


class EpsteinAgent(DeviceAgent):
    """agents.py:13-26."""


class Citizen(EpsteinAgent):
    """agents.py:29-103."""


class Cop(EpsteinAgent):
    """agents.py:106-140."""


def neighborhood_offsets(grid, radius: int) -> np.ndarray:
    """(dx, dy) of Cell._neighborhood(radius) in the reference's order (cell.py:240-276): taken from the grid's own BFS
    (space.neighborhood_coordinates) around a cell, valid for every cell of a torus with both dimensions >= 2 radius + 1."""
    W, H = grid.dimensions
    cx, cy = W // 2, H // 2
    out = []
    for x, y in grid.neighborhood_coordinates((cx, cy), radius):
        dx, dy = (x - cx + W // 2) % W - W // 2, (y - cy + H // 2) % H - H // 2
        if W % 2 == 0 and dx == -W // 2 and radius >= W // 2:
            raise ValueError("grid too narrow for the neighbourhood radius")
        out.append((dx, dy))
    return np.asarray(out, dtype=np.int16).reshape(-1, 2)


@batch_operator(EpsteinAgent, "step")
def _epstein_step(agents: DeviceAgentSet, epoch: int | None = None, **_):
    """`shuffle_do("step")`: the Philox order of `epoch`; `do("step")` (no epoch): registration order, the draws keyed on a
    fresh epoch of their own."""
    m = agents.model
    if epoch is None:
        m._run_activation(m._next_epoch(), sequential=True)
    else:
        m._run_activation(epoch)


class EpsteinCivilViolence(DeviceModel):
    """Model 1 of Epstein, "Modeling civil violence" (model.py:34-125)."""

    def __init__(self, width: int = 40, height: int = 40, scenario: EpsteinScenario = EpsteinScenario, device="cuda",
                 initial_state: dict | None = None):
        if isinstance(scenario, type):
            scenario = scenario()
        super().__init__(scenario=scenario)
        sc = scenario
        if sc.grid_type == "Moore":                                         # model.py:54-67
            self.grid = DeviceMooreGrid((width, height), capacity=1, torus=True, random=self.random, device=device)
        elif sc.grid_type == "Von Neumann":
            self.grid = DeviceVonNeumannGrid((width, height), capacity=1, torus=True, random=self.random, device=device)
        else:
            raise ValueError(f"Unknown value of grid_type: {sc.grid_type}")
        if sc.cop_density + sc.citizen_density > 1:                         # model.py:81-82
            raise ValueError("Cop density + citizen density must be less than 1")
        if sc.activation_order not in ("Random", "Sequential"):
            raise ValueError(f"unknown value of activation_order: {sc.activation_order}")
        vision = int(sc.cop_vision)             # agents.py:18: every agent looks with cop_vision
        if min(width, height) < 2 * vision + 1:
            raise NotImplementedError("the device model needs both grid dimensions >= 2 * vision + 1")
        self.width, self.height = int(width), int(height)
        dev = self.grid.device
        if initial_state is None:
            initial_state = self._reference_initial_state(sc)
        st = initial_state
        kind = torch.as_tensor(np.asarray(st["kind"], dtype=np.uint8)).to(dev).contiguous()
        n = int(kind.numel())
        hardship = torch.as_tensor(np.asarray(st["hardship"], dtype=np.float64)).to(dev).contiguous()
        cols = {
            "cell": torch.as_tensor(np.asarray(st["cell"])).to(torch.int32).to(dev).contiguous(),
            "kind": kind,
            "hardship": hardship,
            "risk_aversion": torch.as_tensor(np.asarray(st["risk_aversion"], dtype=np.float64)).to(dev).contiguous(),
            "grievance": (hardship * (1 - sc.legitimacy)).contiguous(),     # agents.py:60
            "state": torch.where(kind == 0, CitizenState.QUIET.value, 0).to(torch.uint8).contiguous(),
            "jail_sentence": torch.zeros(n, dtype=torch.int32, device=dev),
            "arrest_probability": torch.full((n,), float("nan"), dtype=torch.float64, device=dev),   # None until the first step
        }
        if "state" in st:
            cols["state"] = torch.as_tensor(np.asarray(st["state"], dtype=np.uint8)).to(dev).contiguous()
        if "jail_sentence" in st:
            cols["jail_sentence"] = torch.as_tensor(np.asarray(st["jail_sentence"])).to(torch.int32).to(dev).contiguous()
        agents = DeviceAgentSet(self, EpsteinAgent, n, cols, self.random, resizable=False)
        self._agents = agents
        rows = torch.arange(n, device=dev)
        for klass, k in ((Citizen, 0), (Cop, 1)):
            self._agents_by_type[klass] = DeviceAgentSet(self, klass, n, cols, self.random, rows[kind == k], False, agents._pop)
        self.grid.attach_population(agents)
        # the neighbourhood in the reference's order, the dependency range (2 vision, any order), the probability table
        voff = neighborhood_offsets(self.grid, vision)
        reach = 2 * vision
        if self.grid._moore:
            woff = [(dx, dy) for dx in range(-reach, reach + 1) for dy in range(-reach, reach + 1)]
        else:
            woff = [(dx, dy) for dx in range(-reach, reach + 1) for dy in range(-reach + abs(dx), reach - abs(dx) + 1)]
        # (a narrow torus: offsets beyond half the grid wrap onto cells that are listed anyway)
        woff = sorted({((dx + width // 2) % width - width // 2, (dy + height // 2) % height - height // 2) for dx, dy in woff})
        self._voff = torch.as_tensor(voff).to(dev).contiguous()
        self._woff = torch.as_tensor(np.asarray(woff, dtype=np.int16)).to(dev).contiguous()
        # agents.py:99-103 for every value round(cops / actives) can take: the reference's own expression, evaluated here
        ptab = [1 - math.exp(-1 * sc.arrest_prob_constant * mm) for mm in range(len(voff) + 1)]
        self._ptab = torch.as_tensor(np.asarray(ptab, dtype=np.float64)).to(dev).contiguous()
        self._occ = torch.empty(self.width * self.height, dtype=torch.int32, device=dev)
        nbytes = ctypes.c_size_t(0)
        N.check(N.lib().mb_epstein_workspace_bytes(n, self.width * self.height, ctypes.byref(nbytes)))
        self._ws = torch.empty(max(nbytes.value, 1), dtype=torch.uint8, device=dev)
        self._counts = torch.zeros(4, dtype=torch.int32, device=dev)
        self._counts_host = None
        self._rebuild_occ()
        self.grid.property_layers.hooks["empty"] = self._sync_empty
        self._update_counts()
        self.datacollector = DeviceDataCollector(
            model_reporters={"active": "ACTIVE", "quiet": "QUIET", "arrested": "ARRESTED"},               # model.py:69-73
            agent_reporters={"jail_sentence": "jail_sentence", "arrest_probability": "arrest_probability"})
        self.running = True
        self.datacollector.collect(self)

    # ------------------------------------------------------------------ initial state
    def _reference_initial_state(self, sc) -> dict:
        """model.py:84-98 on the model's own generator (reference_init.epstein_initial_state): the same seed gives the
        reference's agents."""
        from ..reference_init import epstein_initial_state

        return epstein_initial_state(self.random, self.width, self.height, sc.citizen_density, sc.cop_density)

    # ------------------------------------------------------------------ device state
    def _rebuild_occ(self):
        a, dev = self._agents, self.grid.device
        with torch.cuda.device(dev):
            N.check(N.lib().mb_epstein_build_occ(N.ptr(a.columns["cell"]), N.ptr(a.columns["kind"]), N.ptr(a.columns["state"]),
                                                 a.n, N.ptr(self._occ), self.width * self.height, N.stream_ptr(dev)))

    def _sync_empty(self):
        """`grid.empty` / `cell.is_empty` (agents.py:20) from the occupancy words."""
        dict.__getitem__(self.grid.property_layers, "empty").copy_((self._occ < 0).reshape(self.width, self.height))

    def _run_activation(self, epoch: int, sequential: bool = False):
        a, sc, dev = self._agents, self.scenario, self.grid.device
        c = a.columns
        with torch.cuda.device(dev):
            N.check(N.lib().mb_epstein_step(
                N.ptr(self._occ), N.ptr(c["cell"]), N.ptr(c["kind"]), N.ptr(c["state"]), N.ptr(c["jail_sentence"]),
                N.ptr(c["risk_aversion"]), N.ptr(c["grievance"]), N.ptr(c["arrest_probability"]), a.n, self.width, self.height,
                N.ptr(self._voff), int(self._voff.shape[0]), N.ptr(self._woff), int(self._woff.shape[0]), N.ptr(self._ptab),
                int(sc.max_jail_term), float(sc.active_threshold), int(bool(sc.movement)), int(sequential), self.philox_seed, epoch,
                N.ptr(self._counts), N.ptr(self._ws), self._ws.numel(), N.stream_ptr(dev)))
        self._counts_host = None

    # ------------------------------------------------------------------ the reference's model code
    @property
    def agents(self):
        return self._agents

    def _counts_now(self):
        if self._counts_host is None:
            self._counts_host = self._counts.tolist()
            if self._counts_host[0] != 0:
                raise RuntimeError("epstein: a dependency wait ran into its spin limit; the step is not valid")
        return self._counts_host

    ACTIVE = property(lambda self: self._counts_now()[1])
    QUIET = property(lambda self: self._counts_now()[2])
    ARRESTED = property(lambda self: self._counts_now()[3])

    def _update_counts(self):
        """model.py:119-125 `agents_by_type[Citizen].groupby("state").count()`: the step kernel leaves the three counts on
        the device; before the first step they come from the same groupby as the reference's."""
        if self._epoch == 0:
            counts = self.agents_by_type[Citizen].groupby("state").count()
            self._counts_host = [0] + [int(counts.get(s.value, 0)) for s in CitizenState]

    def step(self):
        """model.py:100-117."""
        if self.scenario.activation_order == "Random":
            self.agents.shuffle_do("step")
        else:
            self.agents.do("step")
        self._update_counts()
        self.datacollector.collect(self)
        if self.time > self.scenario.max_iters:
            self.running = False
