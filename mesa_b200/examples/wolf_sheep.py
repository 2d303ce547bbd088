"""Wolf-Sheep predation on the device (mirror of mesa/examples/advanced/wolf_sheep/model.py:22-141, agents.py:4-158):
sheep and wolves walk on a von Neumann torus, sheep eat grass, wolves eat sheep, both die and reproduce.

The reference model's own step runs unchanged --

    self.agents_by_type[Sheep].shuffle_do("step")
    self.agents_by_type[Wolf].shuffle_do("step")

-- on two DYNAMIC device populations (mesa_b200.agentset: stable compaction for the dead, appended rows for the newborn,
ONE id counter for both types like `model.agent_id_counter`), with the grass as a property layer instead of one
GrassPatch agent per cell.  The sequential semantics of each phase are computed exactly by Jacobi passes to a fixed point
(csrc/wolf_sheep.cu); the same seed gives the reference's initial state, and with the replayed Philox order / draws the
reference's trajectory (tests/golden/wolf_sheep_*.npz: ids, cells and float64 energies bit for bit).
"""
from __future__ import annotations

import math

import numpy as np
import torch

from .. import _native as N
from .. import ops
from ..agentset import DeviceAgent, DeviceAgentSet, batch_operator
from ..datacollection import DeviceDataCollector
from ..model import DeviceModel
from ..space import DeviceVonNeumannGrid
from ._scenario import Scenario

_NEVER = -1   # all bits set as int64


class WolfSheepScenario(Scenario):
    """wolf_sheep/model.py:22-50."""

    width: int = 20
    height: int = 20
    initial_sheep: int = 100
    initial_wolves: int = 50
    sheep_reproduce: float = 0.04
    wolf_reproduce: float = 0.05
    wolf_gain_from_food: float = 20.0
    grass: bool = True
    grass_regrowth_time: int = 30
    sheep_gain_from_food: float = 4.0


class Animal(DeviceAgent):
    """agents.py:4-52."""


class Sheep(Animal):
    """agents.py:55-99."""


class Wolf(Animal):
    """agents.py:102-122."""


def _ptr(t):
    return N.ptr(t)


def _fixed_point(pass_fn, size: int, dev, max_passes: int = 100000):
    """Jacobi passes until the table reproduces itself; returns (table, passes).  Tables are uint64 priorities held in int64
    tensors, all bits set = never.  One 4-byte flag is read back per pass."""
    prev = torch.full((max(size, 1),), _NEVER, dtype=torch.int64, device=dev)
    nxt = torch.full((max(size, 1),), _NEVER, dtype=torch.int64, device=dev)
    changed = torch.zeros(1, dtype=torch.int32, device=dev)
    lib, st = N.lib(), N.stream_ptr(dev)
    for k in range(1, max_passes + 1):
        pass_fn(prev, nxt)
        changed.zero_()
        N.check(lib.mb_ws_compare_reset(_ptr(prev), _ptr(nxt), size, _ptr(changed), st))   # prev := never
        prev, nxt = nxt, prev
        if int(changed.item()) == 0:
            return prev, k
    raise RuntimeError("wolf-sheep: the activation did not reach its fixed point")


def _append_offspring(agents: DeviceAgentSet, model, repro, prio, rank, extra):
    """agents.py:24-33 spawn_offspring for all parents of a phase: the offspring are registered in the order in which their
    parents acted (ascending priority), with consecutive ids from the model's counter (model.py:250-263)."""
    parents = ops.nonzero_indices(repro)
    k = int(parents.numel())
    if k == 0:
        return
    by_time = parents[ops.argsort_stable(prio[parents] ^ (-2**63))]       # uint64 order of the parents' priorities
    ids = torch.arange(model.agent_id_counter, model.agent_id_counter + k, dtype=torch.int64, device=parents.device)
    model.agent_id_counter += k
    cols = {name: agents.columns[name][by_time] for name in ("cell", "energy")}
    cols.update({name: fn(by_time) for name, fn in extra.items()})
    agents.add_agents(k, unique_ids=ids, **cols)


@batch_operator(Sheep, "step")
def _sheep_step(agents: DeviceAgentSet, epoch: int, **_):
    m, n = agents.model, agents.n
    if n == 0:
        return
    dev, lib, st = m.grid.device, N.lib(), N.stream_ptr(m.grid.device)
    H, W = m.grid.dimensions
    ncells = H * W
    cell, energy, uid = agents.columns["cell"], agents.columns["energy"], agents.unique_ids().contiguous()
    wolves = m.agents_by_type[Wolf]
    wolves_in = torch.empty(ncells, dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        N.check(lib.mb_ws_count_cells(_ptr(wolves.columns["cell"]), wolves.n, _ptr(wolves_in), ncells, st))
        grown = m.grid.grass_grown.reshape(-1) if m.grass else None
        grown0 = grown.clone() if m.grass else None

        def one_pass(prev, nxt):
            N.check(lib.mb_ws_sheep_pass(_ptr(cell), _ptr(uid), n, H, W, m.philox_seed, epoch, _ptr(wolves_in), _ptr(grown0),
                                         _ptr(prev), _ptr(nxt), st))

        E, m.last_sheep_passes = _fixed_point(one_pass, ncells, dev)
        moved = torch.empty(n, dtype=torch.uint8, device=dev)
        dead, repro = torch.empty_like(moved), torch.empty_like(moved)
        prio = torch.empty(n, dtype=torch.int64, device=dev)
        regrow = m.grid.grass_regrow_at.reshape(-1) if m.grass else None
        N.check(lib.mb_ws_sheep_apply(_ptr(cell), _ptr(energy), _ptr(uid), n, H, W, m.philox_seed, epoch, _ptr(wolves_in),
                                      _ptr(grown0), _ptr(grown), _ptr(regrow), _ptr(E), float(m.sheep_gain),
                                      float(m.sheep_reproduce), int(m.time) + int(m.grass_regrowth_time), _ptr(moved), _ptr(dead),
                                      _ptr(repro), _ptr(prio), st))
    # position in the cell lists (cell.py:143-170: arrivals are appended): sheep that stayed keep their place, the others
    # follow in the order they acted, a lamb right after its parent -- 2^32 + 2 * rank (+ 1 for the lamb)
    order = ops.argsort_stable(prio ^ (-2**63))
    rank = torch.empty(n, dtype=torch.int64, device=dev)
    rank[order] = torch.arange(n, dtype=torch.int64, device=dev)
    arrival = (1 << 32) + 2 * rank
    agents.columns["ord"] = torch.where(moved.bool(), arrival, agents.columns["ord"])
    _append_offspring(agents, m, repro, prio, rank, {"ord": lambda rows: (1 << 32) + 2 * rank[rows] + 1})
    if bool(dead.any().item()):
        dead_full = torch.zeros(agents.n, dtype=torch.bool, device=dev)
        dead_full[:n] = dead.bool()
        agents.remove_agents(dead_full)                      # agents.py:49-50 (stable: the survivors keep their order)


@batch_operator(Wolf, "step")
def _wolf_step(agents: DeviceAgentSet, epoch: int, **_):
    m, n = agents.model, agents.n
    dev, lib, st = m.grid.device, N.lib(), N.stream_ptr(m.grid.device)
    H, W = m.grid.dimensions
    ncells = H * W
    sheep = m.agents_by_type[Sheep]
    ns = sheep.n
    # the cells' sheep lists: rows sorted by (cell, position in the list); the positions are renumbered densely
    key = (sheep.columns["cell"].to(torch.int64) << 34) | sheep.columns["ord"]
    _, slist = ops.sort_pairs(key, torch.arange(ns, dtype=torch.int32, device=dev)) if ns else (None, torch.empty(0, dtype=torch.int32, device=dev))
    sheep.columns["ord"] = torch.empty(ns, dtype=torch.int64, device=dev)
    sheep.columns["ord"][slist.long()] = torch.arange(ns, dtype=torch.int64, device=dev)
    sorted_cell = sheep.columns["cell"][slist.long()].contiguous()
    cstart = torch.empty(ncells + 1, dtype=torch.int32, device=dev)
    if n == 0:
        return
    cell, energy, uid = agents.columns["cell"], agents.columns["energy"], agents.unique_ids().contiguous()
    with torch.cuda.device(dev):
        N.check(lib.mb_ws_segment_starts(_ptr(sorted_cell), ns, ncells, _ptr(cstart), st))

        def one_pass(prev, nxt):
            N.check(lib.mb_ws_wolf_pass(_ptr(cell), _ptr(uid), n, H, W, m.philox_seed, epoch, _ptr(slist), _ptr(cstart),
                                        _ptr(prev), _ptr(nxt), st))

        D, m.last_wolf_passes = _fixed_point(one_pass, ns, dev)
        dead = torch.empty(n, dtype=torch.uint8, device=dev)
        repro = torch.empty_like(dead)
        prio = torch.empty(n, dtype=torch.int64, device=dev)
        N.check(lib.mb_ws_wolf_apply(_ptr(cell), _ptr(energy), _ptr(uid), n, H, W, m.philox_seed, epoch, _ptr(slist), _ptr(cstart),
                                     _ptr(D), float(m.wolf_gain), float(m.wolf_reproduce), _ptr(dead), _ptr(repro), _ptr(prio), st))
    if ns:
        eaten = D[:ns] != _NEVER
        if bool(eaten.any().item()):
            sheep.remove_agents(eaten)                       # agents.py:111 sheep_to_eat.remove()
    _append_offspring(agents, m, repro, prio, None, {})
    if bool(dead.any().item()):
        dead_full = torch.zeros(agents.n, dtype=torch.bool, device=dev)
        dead_full[:n] = dead.bool()
        agents.remove_agents(dead_full)


def _report_wolves(m):
    return len(m.agents_by_type[Wolf])


def _report_sheep(m):
    return len(m.agents_by_type[Sheep])


def _report_grass(m):
    return int(ops.layer_reduce(m.grid.grass_grown, "sum").item())


class WolfSheep(DeviceModel):
    """Wolf-Sheep Predation Model (wolf_sheep/model.py:53-141)."""

    def __init__(self, scenario: WolfSheepScenario = WolfSheepScenario, device="cuda", initial_state: dict | None = None):
        if isinstance(scenario, type):
            scenario = scenario()
        super().__init__(scenario=scenario)
        self.height, self.width, self.grass = scenario.height, scenario.width, bool(scenario.grass)
        self.sheep_reproduce, self.wolf_reproduce = scenario.sheep_reproduce, scenario.wolf_reproduce
        self.sheep_gain, self.wolf_gain = scenario.sheep_gain_from_food, scenario.wolf_gain_from_food
        self.grass_regrowth_time = scenario.grass_regrowth_time
        # model.py:76-81: Grid([height, width], torus=True, capacity=inf)
        self.grid = DeviceVonNeumannGrid((self.height, self.width), torus=True, capacity=math.inf, random=self.random, device=device)
        dev = self.grid.device
        ncells = self.height * self.width
        if initial_state is None:
            initial_state = self._reference_initial_state(scenario, ncells)
        st = initial_state
        self.agent_id_counter = 1
        self.last_sheep_passes = self.last_wolf_passes = 0
        for klass, key in ((Sheep, "sheep"), (Wolf, "wolf")):
            cells = torch.as_tensor(np.asarray(st[f"{key}_cell"])).to(torch.int32).to(dev).contiguous()
            k = int(cells.numel())
            cols = {"cell": cells, "energy": torch.as_tensor(np.asarray(st[f"{key}_energy"], dtype=np.float64)).to(dev).contiguous()}
            if klass is Sheep:
                cols["ord"] = torch.arange(k, dtype=torch.int64, device=dev)     # the cells' lists start in creation order
            aset = DeviceAgentSet(self, klass, k, cols, self.random)
            ids = st.get(f"{key}_id")
            aset._pop.ids = (torch.as_tensor(np.asarray(ids)).to(torch.int64).to(dev) if ids is not None else
                             torch.arange(self.agent_id_counter, self.agent_id_counter + k, dtype=torch.int64, device=dev))
            self.agent_id_counter = max(self.agent_id_counter, int(aset._pop.ids.max().item()) + 1 if k else self.agent_id_counter)
            self.register_agentset(aset)
            self.grid.attach_population(aset)
        if self.grass:
            grown = torch.as_tensor(np.asarray(st["grown"], dtype=np.uint8)).to(dev).reshape(self.height, self.width).contiguous()
            regrow = torch.as_tensor(np.asarray(st["regrow_at"])).to(torch.int32).to(dev).reshape(self.height, self.width).contiguous()
            self.grid.add_property_layer("grass_grown", grown)
            self.grid.add_property_layer("grass_regrow_at", regrow)
            self.agent_id_counter += ncells          # the reference registers one GrassPatch per cell (model.py:120-131)
        self.agent_id_counter = int(st.get("next_id", self.agent_id_counter))
        reporters = {"Wolves": _report_wolves, "Sheep": _report_sheep}       # (module-level functions: models pickle)
        if self.grass:
            reporters["Grass"] = _report_grass
        self.datacollector = DeviceDataCollector(model_reporters=reporters)
        self.running = True
        self.datacollector.collect(self)

    def _reference_initial_state(self, scenario, ncells: int) -> dict:
        """model.py:95-131 with the model's own two generators, call for call: the same seed gives the reference's sheep,
        wolves and grass."""
        st = {}
        for key, n, gain in (("sheep", scenario.initial_sheep, scenario.sheep_gain_from_food),
                             ("wolf", scenario.initial_wolves, scenario.wolf_gain_from_food)):
            st[f"{key}_energy"] = self.rng.random((n,)) * 2 * gain
            st[f"{key}_cell"] = np.asarray(self.random.choices(range(ncells), k=n), dtype=np.int64)
        if scenario.grass:
            grown = np.zeros(ncells, dtype=np.uint8)
            regrow_at = np.full(ncells, -1, dtype=np.int64)
            for c in range(ncells):                        # `for cell in self.grid`: row-major
                fully_grown = self.random.choice([True, False])
                countdown = 0 if fully_grown else self.random.randrange(0, scenario.grass_regrowth_time)
                grown[c] = countdown == 0                  # agents.py:138
                if countdown:
                    regrow_at[c] = countdown               # schedule_event(self.regrow, after=countdown) at time 0
            st["grown"], st["regrow_at"] = grown, regrow_at
        return st

    @property
    def agents(self):
        raise AttributeError("WolfSheep keeps one population per species: use model.agents_by_type[Sheep / Wolf]")

    def step(self):
        """model.py:136-141; the regrowth events that are due fire after the animals have acted (mesa/model.py:156-190)."""
        self.agents_by_type[Sheep].shuffle_do("step")
        self.agents_by_type[Wolf].shuffle_do("step")
        self.datacollector.collect(self)
        if self.grass:
            dev = self.grid.device
            with torch.cuda.device(dev):
                N.check(N.lib().mb_ws_regrow(N.ptr(self.grid.grass_grown), N.ptr(self.grid.grass_regrow_at),
                                             self.height * self.width, int(self.time), N.stream_ptr(dev)))
