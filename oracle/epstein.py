"""Sequential restatement of Epstein's civil-violence model -- TEST INFRASTRUCTURE (oracle; the product never imports it).

Follows mesa/examples/advanced/epstein_civil_violence (agents.py:14-140, model.py:30-125) statement by statement over flat
arrays, with the activation order and the draws of the keyed Philox protocol (oracle/philox.py), i.e. what the unmodified
reference does under oracle/replay.ReplayRandom.  Pinned by tests/golden/epstein_*.npz (oracle/make_golden.py runs the
live reference) and by the live differential test.

    cell index     flat = x * height + y                      (grid dimensions are (width, height), model.py:56-63)
    neighbourhood  Cell._neighborhood(radius = cop_vision): BFS over the connections (cell.py:240-276), centre excluded,
                   in BFS order -- `empty_neighbors` and `neighbors` keep that order (agents.py:18-20)
    citizen        agents.py:66-83  (jailed: count down and return; else arrest probability, state, move)
    cop            agents.py:124-140 (arrest a random active citizen in vision, then move)
    draws          random.choice(seq) = seq[_randbelow(len(seq))], randint(0, J) = _randbelow(J + 1)
"""
from __future__ import annotations

import math

import numpy as np

from . import philox

QUIET, ACTIVE, ARRESTED = 2, 1, 3          # CitizenState values (agents.py:7-10)
CITIZEN, COP = 0, 1

MOORE = [(-1, -1), (-1, 0), (-1, 1), (0, -1), (0, 1), (1, -1), (1, 0), (1, 1)]      # grid.py:447-451
VON_NEUMANN = [(-1, 0), (0, -1), (0, 1), (1, 0)]                                     # grid.py:478-482


def neighborhood_offsets(moore: bool, radius: int) -> list[tuple[int, int]]:
    """Offsets of Cell._neighborhood(radius) in the reference's BFS order on a torus that is wide enough not to wrap onto
    itself (both dimensions >= 2 * radius + 1): cell.py:240-276."""
    conn = MOORE if moore else VON_NEUMANN
    if radius == 1:
        return list(conn)
    visited = {(0, 0)}
    layer = list(conn)
    found = {}
    for c in layer:
        if c not in visited:
            visited.add(c)
            found[c] = None
    for _ in range(radius - 1):
        nxt = []
        for c in layer:
            for d in conn:
                nb = (c[0] + d[0], c[1] + d[1])
                if nb not in visited:
                    visited.add(nb)
                    nxt.append(nb)
                    found[nb] = None
        layer = nxt
    found.pop((0, 0), None)
    return list(found)


class _Draws:
    """The keyed draw stream of one activation (replay.ReplayRandom.getrandbits)."""

    def __init__(self, seed, epoch, agent):
        self.seed, self.epoch, self.agent, self.c = seed, epoch, agent, 0

    def randbelow(self, n: int) -> int:
        k = int(n).bit_length()
        while True:
            r = philox.getrandbits(self.seed, self.epoch, self.agent, self.c, k)
            self.c += 1
            if r < n:
                return r


class EpsteinOracle:
    """State: kind / cell / state / jail / hardship / risk_aversion / arrest_probability per agent (index = unique_id - 1),
    occ[cell] = agent index or -1."""

    def __init__(self, width, height, kind, cell, hardship, risk_aversion, *, moore=False, vision=7, legitimacy=0.8,
                 max_jail_term=1000, active_threshold=0.1, arrest_prob_constant=2.3, movement=True, philox_seed=0,
                 state=None, jail=None):
        self.width, self.height = int(width), int(height)
        if min(self.width, self.height) < 2 * vision + 1:
            raise ValueError("the oracle's offset tables need both dimensions >= 2 * vision + 1")
        self.kind = np.asarray(kind, dtype=np.int8).copy()
        self.cell = np.asarray(cell, dtype=np.int64).copy()
        n = len(self.kind)
        self.hardship = np.asarray(hardship, dtype=np.float64).copy()
        self.risk_aversion = np.asarray(risk_aversion, dtype=np.float64).copy()
        self.grievance = self.hardship * (1 - legitimacy)                       # agents.py:60
        self.state = np.full(n, QUIET, dtype=np.int8) if state is None else np.asarray(state, dtype=np.int8).copy()
        self.jail = np.zeros(n, dtype=np.int64) if jail is None else np.asarray(jail, dtype=np.int64).copy()
        self.arrest_probability = np.full(n, np.nan)                            # None until the first step
        self.offsets = neighborhood_offsets(moore, vision)
        self.max_jail_term, self.threshold, self.k = int(max_jail_term), float(active_threshold), float(arrest_prob_constant)
        self.movement = bool(movement)
        self.seed, self.epoch = int(philox_seed), 0
        self.occ = np.full(self.width * self.height, -1, dtype=np.int64)
        self.occ[self.cell] = np.arange(n)

    def _neighbors(self, c):
        x, y = divmod(int(c), self.height)
        return [((x + dx) % self.width) * self.height + (y + dy) % self.height for dx, dy in self.offsets]

    def step(self):
        n = len(self.kind)
        order = philox.permutation(self.seed, self.epoch, n)
        for a in order:
            a = int(a)
            d = _Draws(self.seed, self.epoch, a)
            if self.kind[a] == CITIZEN:
                if self.jail[a]:
                    self.jail[a] -= 1
                    continue
                nb = self._neighbors(self.cell[a])
                empty = [c for c in nb if self.occ[c] < 0]
                cops, actives = 0, 1
                for c in nb:
                    o = self.occ[c]
                    if o < 0:
                        continue
                    if self.kind[o] == COP:
                        cops += 1
                    elif self.state[o] == ACTIVE:
                        actives += 1
                p = 1 - math.exp(-1 * self.k * round(cops / actives))            # agents.py:99-103
                self.arrest_probability[a] = p
                net_risk = self.risk_aversion[a] * p
                self.state[a] = ACTIVE if (self.grievance[a] - net_risk) > self.threshold else QUIET
            else:
                nb = self._neighbors(self.cell[a])
                empty = [c for c in nb if self.occ[c] < 0]
                active = [int(self.occ[c]) for c in nb
                          if self.occ[c] >= 0 and self.kind[self.occ[c]] == CITIZEN and self.state[self.occ[c]] == ACTIVE]
                if active:
                    arrestee = active[d.randbelow(len(active))]
                    self.jail[arrestee] = d.randbelow(self.max_jail_term + 1)    # randint(0, max_jail_term)
                    self.state[arrestee] = ARRESTED
            if self.movement and empty:
                dest = empty[d.randbelow(len(empty))]
                self.occ[self.cell[a]] = -1
                self.occ[dest] = a
                self.cell[a] = dest
        self.epoch += 1

    def counts(self):
        cit = self.kind == CITIZEN
        return (int((self.state[cit] == ACTIVE).sum()), int((self.state[cit] == QUIET).sum()),
                int((self.state[cit] == ARRESTED).sum()))
