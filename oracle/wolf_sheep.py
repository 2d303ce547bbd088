"""CPU restatement of the Wolf-Sheep model's step (TEST INFRASTRUCTURE: only tests/ and bench.py's CPU legs import it).

Follows, line for line and sequentially,
    mesa/examples/advanced/wolf_sheep/agents.py:36-52   Animal.step (move, energy -= 1, feed, die or reproduce)
    agents.py:58-99                                      Sheep.feed / Sheep.move
    agents.py:105-122                                    Wolf.feed / Wolf.move
    agents.py:125-158                                    GrassPatch (regrowth as a scheduled event)
    mesa/examples/advanced/wolf_sheep/model.py:136-141   step: agents_by_type[Sheep].shuffle_do("step"), then the wolves
with the keyed Philox draws of oracle/philox.py in place of the shared MT19937 stream (the reference is driven with the same
numbers by oracle/replay.py): the activation order of shuffle number `epoch` is ascending priority64(seed, epoch,
unique_id - 1), and the draws of an agent's step are getrandbits / random() number 0, 1, 2, ... of its own stream.

Pinned against the live reference by tests/golden/wolf_sheep_*.npz (oracle/make_golden.py wolf_sheep_fixture) and by the
live differential test in tests/test_live_reference.py.
"""
from __future__ import annotations

import numpy as np

from . import philox

VN_OFFSETS = ((-1, 0), (0, -1), (0, 1), (1, 0))   # grid.py:478-482, connection order


class _Stream:
    """The getrandbits / random() sequence of one activation (CPython random.py:242-250 for _randbelow)."""

    def __init__(self, seed, epoch, agent):
        self.seed, self.epoch, self.agent, self.c = seed, epoch, agent, 0

    def randbelow(self, n: int) -> int:
        k = int(n).bit_length()
        while True:
            r = philox.getrandbits(self.seed, self.epoch, self.agent, self.c, k)
            self.c += 1
            if r < n:
                return r

    def random(self) -> float:
        v = float(philox.uniform01(self.seed, self.epoch, self.agent, self.c))
        self.c += 1
        return v


class WolfSheepState:
    """Agents are dict entries keyed by unique_id (insertion order = registration order, like model._agents_by_type)."""

    def __init__(self, height, width, sheep, wolves, grown, regrow_at, *, grass=True, grass_regrowth_time=30,
                 sheep_reproduce=0.04, wolf_reproduce=0.05, sheep_gain=4.0, wolf_gain=20.0, next_id=None, seed=0):
        """sheep / wolves: iterables of (unique_id, flat cell, energy) in registration order; grown: bool [H, W];
        regrow_at: int [H, W], the time at which an eaten patch regrows (-1: none pending)."""
        self.h, self.w = int(height), int(width)
        self.grass, self.regrowth = bool(grass), int(grass_regrowth_time)
        self.p = {"sheep": float(sheep_reproduce), "wolf": float(wolf_reproduce)}
        self.gain = {"sheep": float(sheep_gain), "wolf": float(wolf_gain)}
        self.seed = int(seed)
        self.sheep = {int(u): [int(c), float(e)] for u, c, e in sheep}
        self.wolves = {int(u): [int(c), float(e)] for u, c, e in wolves}
        self.grown = np.array(grown, dtype=bool).reshape(self.h, self.w).copy()
        self.regrow_at = np.array(regrow_at, dtype=np.int64).reshape(self.h, self.w).copy()
        # per-cell lists of the sheep in arrival order (cell.py:143-170: append on arrival, remove on departure)
        self.cell_sheep = [[] for _ in range(self.h * self.w)]
        for u, (c, _) in self.sheep.items():
            self.cell_sheep[c].append(u)
        ids = list(self.sheep) + list(self.wolves)
        self.next_id = int(next_id) if next_id is not None else (max(ids) + 1 if ids else 1)
        self.time = 0
        self.epoch = 0

    def neighbors(self, c):
        i, j = divmod(c, self.w)
        return [((i + di) % self.h) * self.w + (j + dj) % self.w for di, dj in VN_OFFSETS]

    def step(self):
        """model.py:136-141 at time `self.time + 1`, followed by the regrowth events of that time (mesa/model.py:156-190:
        `step()` advances the clock by one unit and runs every event due; the step itself has the higher priority, so the
        patches regrow AFTER the animals of that step have acted)."""
        self.time += 1
        wolves_in = np.zeros(self.h * self.w, dtype=np.int64)
        for c, _ in self.wolves.values():
            wolves_in[c] += 1
        # ---- sheep, in the order of shuffle number `epoch`
        epoch, self.epoch = self.epoch, self.epoch + 1
        order = sorted(self.sheep, key=lambda u: int(philox.priority64(self.seed, epoch, u - 1)))
        for u in order:
            rs = _Stream(self.seed, epoch, u - 1)
            c, e = self.sheep[u]
            # agents.py:69-99 move
            safe, grassy = [], []
            for n in self.neighbors(c):
                if wolves_in[n] > 0:
                    continue
                safe.append(n)
                if self.grass and self.grown.reshape(-1)[n]:
                    grassy.append(n)
            if safe:
                targets = grassy if grassy else safe
                d = targets[rs.randbelow(len(targets))]
                self.cell_sheep[c].remove(u)
                self.cell_sheep[d].append(u)
                c = d
            e -= 1
            # agents.py:61-67 feed
            if self.grass and self.grown.reshape(-1)[c]:
                e += self.gain["sheep"]
                self.grown.reshape(-1)[c] = False
                self.regrow_at.reshape(-1)[c] = self.time + self.regrowth
            self.sheep[u] = [c, e]
            if e < 0:
                del self.sheep[u]
                self.cell_sheep[c].remove(u)
            elif rs.random() < self.p["sheep"]:
                e /= 2
                self.sheep[u][1] = e
                self.sheep[self.next_id] = [c, e]
                self.cell_sheep[c].append(self.next_id)
                self.next_id += 1
        # ---- wolves
        epoch, self.epoch = self.epoch, self.epoch + 1
        order = sorted(self.wolves, key=lambda u: int(philox.priority64(self.seed, epoch, u - 1)))
        for u in order:
            rs = _Stream(self.seed, epoch, u - 1)
            c, e = self.wolves[u]
            nbrs = self.neighbors(c)
            with_sheep = [n for n in nbrs if self.cell_sheep[n]]
            targets = with_sheep if with_sheep else nbrs
            c = targets[rs.randbelow(len(targets))]
            e -= 1
            here = self.cell_sheep[c]
            if here:
                victim = here[rs.randbelow(len(here))]
                e += self.gain["wolf"]
                here.remove(victim)
                del self.sheep[victim]
            self.wolves[u] = [c, e]
            if e < 0:
                del self.wolves[u]
            elif rs.random() < self.p["wolf"]:
                e /= 2
                self.wolves[u][1] = e
                self.wolves[self.next_id] = [c, e]
                self.next_id += 1

        if self.grass:
            due = (self.regrow_at >= 0) & (self.regrow_at <= self.time)
            self.grown[due] = True
            self.regrow_at[due] = -1

    # ---- views for the tests
    def table(self, kind: str):
        d = self.sheep if kind == "sheep" else self.wolves
        ids = np.array(list(d), dtype=np.int64)
        cells = np.array([v[0] for v in d.values()], dtype=np.int64)
        energy = np.array([v[1] for v in d.values()], dtype=np.float64)
        return ids, cells, energy
