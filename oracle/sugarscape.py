"""Sequential restatement of the Sugarscape G1MT model (movement and trade) -- TEST INFRASTRUCTURE (oracle; the product never
imports it).

Follows mesa/examples/advanced/sugarscape_g1mt (model.py:118-127 without trade, agents.py:209-288) statement by statement
over flat arrays, with the activation order and the draws of the keyed Philox protocol (oracle/philox.py), i.e. what the
unmodified reference does under oracle/replay.ReplayRandom with `enable_trade=False`.  Pinned by
tests/golden/sugarscape_*.npz (oracle/make_golden.py runs the live reference).

    cell index     flat = x * height + y                     (grid dimensions (width, height), von Neumann, NOT a torus)
    regrowth       layer = minimum(layer + 1, distribution)                                         model.py:123-124
    move           empty cells of Cell.get_neighborhood(vision, include_center=True) in BFS order (cell.py:240-276; on a
                   clamped von Neumann grid = the interior order with the absent cells dropped -- checked against the live
                   grid), welfare sugar ** (ms / mt) * spice ** (mp / mt) of (own + cell) amounts, the cells within
                   math.isclose of the best welfare, of those the ones within 1 % of the smallest Euclidean distance,
                   random.choice of these                                                           agents.py:209-262
    eat / die      agents.py:264-281 (the cell's sugar and spice are taken whole; starved = sugar <= 0 or spice <= 0)
    trade          model.py:128-129 a second shuffle; agents.py:290-301: every trader in the radius-vision neighbourhood (BFS
                   order of the cells, centre excluded; a cell's traders in arrival order = ascending unique_id, since traders
                   only ever move to EMPTY cells) is traded with until no trade happens (agents.py:156-207, 113-153)
    keys           priority / draws keyed on unique_id - 1 (agents die: rows are not ids)
"""
from __future__ import annotations

import math

import numpy as np

from . import philox
from .epstein import VON_NEUMANN, _Draws, neighborhood_offsets


class SugarscapeOracle:
    def __init__(self, width, height, sugar_max, spice_max, uid, cell, sugar, spice, metabolism_sugar, metabolism_spice, vision,
                 philox_seed=0, sugar_layer=None, spice_layer=None):
        self.width, self.height = int(width), int(height)
        self.sugar_max = np.asarray(sugar_max, dtype=np.float64).reshape(-1).copy()
        self.spice_max = np.asarray(spice_max, dtype=np.float64).reshape(-1).copy()
        self.sugar_layer = self.sugar_max.copy() if sugar_layer is None else np.asarray(sugar_layer, dtype=np.float64).reshape(-1).copy()
        self.spice_layer = self.spice_max.copy() if spice_layer is None else np.asarray(spice_layer, dtype=np.float64).reshape(-1).copy()
        self.uid = np.asarray(uid, dtype=np.int64).copy()
        self.cell = np.asarray(cell, dtype=np.int64).copy()
        self.sugar = np.asarray(sugar, dtype=np.float64).copy()
        self.spice = np.asarray(spice, dtype=np.float64).copy()
        self.ms = np.asarray(metabolism_sugar, dtype=np.int64).copy()
        self.mp = np.asarray(metabolism_spice, dtype=np.int64).copy()
        self.vision = np.asarray(vision, dtype=np.int64).copy()
        self.seed, self.epoch = int(philox_seed), 0
        self.offsets = {v: neighborhood_offsets(False, v) + [(0, 0)] for v in sorted(set(self.vision.tolist()))}
        self.count = np.zeros(self.width * self.height, dtype=np.int64)
        np.add.at(self.count, self.cell, 1)

    def welfare(self, a, sugar, spice):
        mt = self.ms[a] + self.mp[a]
        return np.float64(sugar) ** (self.ms[a] / mt) * np.float64(spice) ** (self.mp[a] / mt)      # agents.py:66-70

    def step(self):
        # model.py:123-124
        self.sugar_layer = np.minimum(self.sugar_layer + 1, self.sugar_max)
        self.spice_layer = np.minimum(self.spice_layer + 1, self.spice_max)
        n = len(self.uid)
        pri = philox.priority64(self.seed, self.epoch, (self.uid - 1).astype(np.uint64))
        order = np.argsort(pri, kind="stable")
        dead = np.zeros(n, dtype=bool)
        for a in order:
            a = int(a)
            d = _Draws(self.seed, self.epoch, int(self.uid[a]) - 1)
            x, y = divmod(int(self.cell[a]), self.height)
            cand = []
            for dx, dy in self.offsets[int(self.vision[a])]:
                nx, ny = x + dx, y + dy
                if 0 <= nx < self.width and 0 <= ny < self.height and self.count[nx * self.height + ny] == 0:
                    cand.append((nx * self.height + ny, dx, dy))
            if cand:
                welfares = [self.welfare(a, self.sugar[a] + self.sugar_layer[c], self.spice[a] + self.spice_layer[c]) for c, _, _ in cand]
                best = max(welfares)
                close = [cd for cd, w in zip(cand, welfares) if math.isclose(w, best)]
                dist = [math.sqrt(dx ** 2 + dy ** 2) for _, dx, dy in close]
                dmin = min(dist)
                final = [cd[0] for cd, ds in zip(close, dist) if math.isclose(ds, dmin, rel_tol=1e-02)]
                dest = final[d.randbelow(len(final))]
                self.count[self.cell[a]] -= 1
                self.count[dest] += 1
                self.cell[a] = dest
            c = int(self.cell[a])
            self.sugar[a] += self.sugar_layer[c]; self.sugar_layer[c] = 0; self.sugar[a] -= self.ms[a]     # agents.py:264-271
            self.spice[a] += self.spice_layer[c]; self.spice_layer[c] = 0; self.spice[a] -= self.mp[a]
            if self.sugar[a] <= 0 or self.spice[a] <= 0:                                                     # agents.py:273-281
                dead[a] = True
                self.count[c] -= 1
        keep = ~dead
        for name in ("uid", "cell", "sugar", "spice", "ms", "mp", "vision"):
            setattr(self, name, getattr(self, name)[keep])
        self.epoch += 1

    # ------------------------------------------------------------------ trade (agents.py:84-207, 290-301)
    def mrs(self, a, sugar, spice):
        return (spice / self.mp[a]) / (sugar / self.ms[a])                                       # agents.py:84-92

    def _maybe_sell_spice(self, seller, buyer, price, w_seller, w_buyer):
        if price >= 1:                                                                           # agents.py:94-106
            sugar_ex, spice_ex = 1, int(price)
        else:
            sugar_ex, spice_ex = int(1 / price), 1
        s_sugar, b_sugar = self.sugar[seller] + sugar_ex, self.sugar[buyer] - sugar_ex
        s_spice, b_spice = self.spice[seller] - spice_ex, self.spice[buyer] + spice_ex
        if s_sugar <= 0 or b_sugar <= 0 or s_spice <= 0 or b_spice <= 0:
            return False
        better = (w_seller < self.welfare(seller, s_sugar, s_spice)) and (w_buyer < self.welfare(buyer, b_sugar, b_spice))
        not_crossing = self.mrs(seller, s_sugar, s_spice) > self.mrs(buyer, b_sugar, b_spice)
        if not (better and not_crossing):
            return False
        self.sugar[seller], self.sugar[buyer], self.spice[seller], self.spice[buyer] = s_sugar, b_sugar, s_spice, b_spice
        return True

    def trade_phase(self):
        """model.py:128-129; returns per trader (rows) the lists of prices and partner ids of this step."""
        n = len(self.uid)
        pri = philox.priority64(self.seed, self.epoch, (self.uid - 1).astype(np.uint64))
        order = np.argsort(pri, kind="stable")
        by_cell = {}
        for i in np.argsort(self.uid, kind="stable"):
            by_cell.setdefault(int(self.cell[i]), []).append(int(i))
        prices, partners = [[] for _ in range(n)], [[] for _ in range(n)]
        for a in order:
            a = int(a)
            x, y = divmod(int(self.cell[a]), self.height)
            for dx, dy in self.offsets[int(self.vision[a])][:-1]:
                nx, ny = x + dx, y + dy
                if not (0 <= nx < self.width and 0 <= ny < self.height):
                    continue
                for b in by_cell.get(nx * self.height + ny, ()):
                    while True:                                                                  # agents.py:156-207 (the recursion)
                        mrs_a, mrs_b = self.mrs(a, self.sugar[a], self.spice[a]), self.mrs(b, self.sugar[b], self.spice[b])
                        w_a, w_b = self.welfare(a, self.sugar[a], self.spice[a]), self.welfare(b, self.sugar[b], self.spice[b])
                        if math.isclose(mrs_a, mrs_b):
                            break
                        price = math.sqrt(mrs_a * mrs_b)
                        sold = (self._maybe_sell_spice(a, b, price, w_a, w_b) if mrs_a > mrs_b else
                                self._maybe_sell_spice(b, a, price, w_b, w_a))
                        if not sold:
                            break
                        prices[a].append(price)
                        partners[a].append(int(self.uid[b]))
        self.epoch += 1
        return prices, partners
