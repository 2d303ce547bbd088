"""Seeds and initial states exactly as the reference derives them, vectorised.

A model built here from `rng=42` starts from the SAME state as the reference model built from `rng=42`:

* seeding -- `Scenario.__init__` (mesa/experimental/scenarios/scenario.py:156-158): `rng = np.random.default_rng(rng)`;
  `Scenario._stdlib_seed` (:214-221): the first uint32 of `seed_seq.generate_state(1)` seeds `model.random =
  random.Random(...)` (mesa/model.py:125-129);
* the example models then draw their initial state cell by cell / agent by agent from `model.random`
  (schelling/model.py:71-81, conways_game_of_life/model.py:19-27, boltzmann_wealth_model/model.py:70-74).

`random.random()` is MT19937's genrand_res53; numpy's legacy `RandomState.random_sample` on an `MT19937` bit generator
holding the same 624-word state produces the same doubles, so one vectorised draw replaces the per-cell Python loop, and
the Python generator is advanced past exactly the numbers the reference would have consumed.  Host-side set-up code:
nothing here is on the per-step path.
"""
from __future__ import annotations

import random as _random

import numpy as np

# above this many uniforms the models fall back to their numpy-Generator initialisation (the reference cannot build
# grids of that size at all: ~1.3 KB and 14 us per cell)
MAX_REFERENCE_DRAWS = 1 << 26


def seed_streams(rng=None):
    """(numpy Generator, stdlib seed) like Scenario(rng=rng): scenario.py:156-158, 214-221."""
    gen = np.random.default_rng(rng)
    seed_seq = gen.bit_generator.seed_seq
    return gen, int(seed_seq.generate_state(1, dtype=np.uint32)[0])


def _bitgen_from(rnd: _random.Random):
    version, state, gauss = rnd.getstate()
    if version != 3:
        raise RuntimeError("unexpected random.Random state version")
    bg = np.random.MT19937()
    bg.state = {"bit_generator": "MT19937", "state": {"key": np.array(state[:-1], dtype=np.uint32), "pos": int(state[-1])}}
    return bg, version, gauss


def mt_uniforms(rnd: _random.Random, n: int, advance: bool = True) -> np.ndarray:
    """The next `n` values of `rnd.random()` in one vectorised draw; `advance` moves `rnd` past them."""
    bg, version, gauss = _bitgen_from(rnd)
    out = np.random.RandomState(bg).random_sample(int(n))
    if advance:
        st = bg.state["state"]
        rnd.setstate((version, tuple(int(x) for x in st["key"]) + (int(st["pos"]),), gauss))
    return out


def gol_initial_state(rnd: _random.Random, width: int, height: int, fraction_alive: float) -> np.ndarray:
    """conways_game_of_life/model.py:19-27: row-major over the cells, alive iff `random() < fraction`."""
    return (mt_uniforms(rnd, width * height) < fraction_alive).reshape(width, height)


def boltzmann_initial_cells(rnd: _random.Random, n: int, ncells: int) -> np.ndarray:
    """boltzmann_wealth_model/model.py:70-74 `random.choices(cells, k=n)`: CPython takes
    `population[floor(random() * len(population))]` for each of the k picks (random.py `choices`, no weights)."""
    return np.floor(mt_uniforms(rnd, n) * float(ncells)).astype(np.int64)


def schelling_initial_types(rnd: _random.Random, width: int, height: int, density: float, minority_pc: float) -> np.ndarray:
    """schelling/model.py:71-81: row-major over the cells, `random() < density` makes an agent, and only then a second
    draw `random() < minority_pc` picks type 1.  int8 [width, height]: -1 empty, 0 / 1 the type.

    The stream position of a cell's first draw depends on how many earlier cells were occupied.  With o[i] =
    (u[i] < density), index i is a TYPE draw iff i-1 was an occupancy draw with o[i-1]: s[i] = (1 - s[i-1]) * o[i-1].  A
    0 in o resets s, inside a run of 1s it alternates, so s[i] = (i - start of the run of 1s that ends at i-1) mod 2 --
    a running maximum instead of a loop."""
    ncells = int(width) * int(height)
    u = mt_uniforms(rnd, 2 * ncells, advance=False)          # at most two draws per cell
    o = u < density
    idx = np.arange(2 * ncells, dtype=np.int64)
    # run_start[i] = first index of the run of ones covering i-1 (i itself if o[i-1] is 0)
    prev_one = np.concatenate([[False], o[:-1]])
    last_reset = np.maximum.accumulate(np.where(prev_one, 0, idx))   # latest j <= i with o[j-1] == 0 (j = 0 counts)
    is_type_draw = ((idx - last_reset) & 1).astype(bool)
    occ_draws = np.flatnonzero(~is_type_draw)[:ncells]       # the first draw of cell 0, 1, ...
    occupied = o[occ_draws]
    types = np.full(ncells, -1, dtype=np.int8)
    types[occupied] = (u[occ_draws[occupied] + 1] < minority_pc).astype(np.int8)
    mt_uniforms(rnd, ncells + int(occupied.sum()))           # advance past exactly what the reference consumed
    return types.reshape(width, height)


def epstein_initial_state(rnd: _random.Random, width: int, height: int, citizen_density: float, cop_density: float) -> dict:
    """epstein_civil_violence/model.py:84-98 with the model's own generator, call for call: `random.choices([Citizen, Cop,
    None], cum_weights=...)` per cell in `all_cells` order (x-major) and, for a citizen, the two `random.random()` of
    agents.py:55-56 (hardship, risk_aversion).  Agents in creation order (unique_id - 1 = row): kind 0 citizen / 1 cop."""
    cum = [citizen_density, citizen_density + cop_density, 1]
    kind, cell, hardship, risk = [], [], [], []
    for c in range(int(width) * int(height)):
        k = rnd.choices([0, 1, None], cum_weights=cum)[0]
        if k is None:
            continue
        kind.append(k)
        cell.append(c)
        hardship.append(rnd.random() if k == 0 else 0.0)
        risk.append(rnd.random() if k == 0 else 0.0)
    return {"kind": np.asarray(kind, dtype=np.uint8), "cell": np.asarray(cell, dtype=np.int64),
            "hardship": np.asarray(hardship, dtype=np.float64), "risk_aversion": np.asarray(risk, dtype=np.float64)}


def sugarscape_initial_state(rnd: _random.Random, gen: np.random.Generator, ncells: int, initial_population: int, endowment_min: int,
                             endowment_max: int, metabolism_min: int, metabolism_max: int, vision_min: int, vision_max: int) -> dict:
    """sugarscape_g1mt/model.py:80-116 on the model's two generators, call for call: the cells from `random.choices(cells, k=n)`
    (stdlib stream), then sugar, spice, metabolism_sugar, metabolism_spice and vision from `rng.integers(lo, hi, (n,),
    endpoint=True)` (numpy stream) in the order the arguments of `Trader.create_agents` are evaluated."""
    n = int(initial_population)
    cells = np.asarray(rnd.choices(range(int(ncells)), k=n), dtype=np.int64)
    draw = lambda lo, hi: gen.integers(lo, hi, (n,), endpoint=True)       # noqa: E731
    return {"cell": cells, "sugar": draw(endowment_min, endowment_max), "spice": draw(endowment_min, endowment_max),
            "metabolism_sugar": draw(metabolism_min, metabolism_max), "metabolism_spice": draw(metabolism_min, metabolism_max),
            "vision": draw(vision_min, vision_max)}
