"""Generate tests/golden/*.npz from the LIVE reference -- run in the build container only.

    python -m oracle.make_golden

Imports the unmodified reference from /root/reference, runs the four example models (the
Schelling / Boltzmann activations replayed with the keyed Philox draws, oracle/replay.py),
and stores small input/output vectors.  The fixtures pin the CPU oracle
(tests/test_oracle_golden.py) and the CUDA path (tests/test_*_gpu.py); they travel to the GPU
box, the reference does not.
"""
from __future__ import annotations

import hashlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
GOLDEN = os.path.join(ROOT, "tests", "golden")
sys.path.insert(0, ROOT)

from oracle.replay import ReplayRandom, activation_context, import_reference, patched_model_random  # noqa: E402


def digest(a: np.ndarray) -> np.ndarray:
    return np.frombuffer(hashlib.sha256(np.ascontiguousarray(a).tobytes()).digest()[:8], dtype=np.uint64)[0]


# ------------------------------------------------------------------------------ Game of Life
def gol_fixture(width, height, seed, steps, name):
    from mesa.examples.basic.conways_game_of_life.model import ConwaysGameOfLife

    m = ConwaysGameOfLife(width=width, height=height, rng=seed)

    def state():
        s = np.zeros((width, height), np.uint8)
        for a in m.agents:
            s[a.cell.coordinate] = a.state
        return s

    states = [state()]
    for _ in range(steps):
        m.step()
        states.append(state())
    np.savez_compressed(os.path.join(GOLDEN, name), states=np.stack(states), torus=np.array(True))
    print(name, np.stack(states).shape)


# ------------------------------------------------------------------------------ Schelling
def schelling_fixture(name, width, height, density, homophily, radius, steps, mt_seed, philox_seed,
                      full_steps):
    from mesa.examples.basic.schelling.agents import SchellingAgent
    from mesa.examples.basic.schelling.model import Schelling, SchellingScenario

    with patched_model_random(philox_seed):
        m = Schelling(SchellingScenario(width=width, height=height, density=density, minority_pc=0.5,
                                        homophily=homophily, radius=radius, rng=mt_seed))
    rng = m.random
    assert isinstance(rng, ReplayRandom) and m.grid.random is rng
    agents = sorted(m.agents, key=lambda a: a.unique_id)
    assert [a.unique_id for a in agents] == list(range(1, len(agents) + 1))

    def cells():
        return np.array([a.cell.coordinate[0] * height + a.cell.coordinate[1] for a in agents], dtype=np.uint32)

    def happys():
        return np.array([a.happy for a in agents], dtype=np.uint8)

    out = {
        "width": width, "height": height, "homophily": homophily, "radius": radius,
        "philox_seed": np.uint64(philox_seed),
        "atype": np.array([a.type for a in agents], dtype=np.int8),
        "cell0": cells(), "happy0": happys(), "happy_count0": m.happy,
    }
    happy_counts, movers, cell_digest, happy_digest = [], [], [], []
    rng.replay = True
    with activation_context(SchellingAgent, "step"):
        for s in range(1, steps + 1):
            before = cells()
            m.step()
            after = cells()
            happy_counts.append(m.happy)
            movers.append(int((before != after).sum()))
            cell_digest.append(digest(after))
            happy_digest.append(digest(happys()))
            if s in full_steps:
                out[f"cell_{s}"] = after
                out[f"happy_{s}"] = happys()
            assert np.array_equal(m.grid.empty.reshape(-1), np.bincount(after, minlength=width * height) == 0)
    rng.replay = False
    out.update(happy_count=np.array(happy_counts), movers=np.array(movers),
               cell_digest=np.array(cell_digest, dtype=np.uint64),
               happy_digest=np.array(happy_digest, dtype=np.uint64), full_steps=np.array(sorted(full_steps)))
    np.savez_compressed(os.path.join(GOLDEN, name), **out)
    print(name, "agents", len(agents), "happy", happy_counts[:5], "...", happy_counts[-1], "movers", movers[:5])


def datacollector_fixture(name, width, height, density, homophily, radius, steps, mt_seed, philox_seed):
    """The UNMODIFIED reference Schelling model's own DataCollector output (model.py:53-70, collected at construction
    and after every step) under the replayed activation order: model-level reporters and the (Step, AgentID)-indexed
    agent frame.  Pins mesa_b200.datacollection.DeviceDataCollector."""
    from mesa.examples.basic.schelling.agents import SchellingAgent
    from mesa.examples.basic.schelling.model import Schelling, SchellingScenario

    with patched_model_random(philox_seed):
        m = Schelling(SchellingScenario(width=width, height=height, density=density, minority_pc=0.5,
                                        homophily=homophily, radius=radius, rng=mt_seed))
    agents = sorted(m.agents, key=lambda a: a.unique_id)
    types = np.full((width, height), -1, dtype=np.int8)
    for a in agents:
        types[a.cell.coordinate] = a.type
    m.random.replay = True
    with activation_context(SchellingAgent, "step"):
        for _ in range(steps):
            m.step()
    m.random.replay = False
    mv = m.datacollector.get_model_vars_dataframe()
    av = m.datacollector.get_agent_vars_dataframe().reset_index()
    out = {"width": width, "height": height, "density": density, "homophily": homophily, "radius": radius, "steps": steps,
           "philox_seed": np.uint64(philox_seed), "types0": types,
           "model_columns": np.array(list(mv.columns)), "agent_step": av["Step"].to_numpy(dtype=np.float64),
           "agent_id": av["AgentID"].to_numpy(dtype=np.int64), "agent_type": av["agent_type"].to_numpy(dtype=np.int64)}
    for c in mv.columns:
        out["model_" + c] = mv[c].to_numpy(dtype=np.float64)
    np.savez_compressed(os.path.join(GOLDEN, name), **out)
    print(name, "rows", len(mv), "agent records", len(av), "happy", list(mv["happy"])[:4])


# ------------------------------------------------------------------------------ neighbourhoods
def neighborhood_fixture(name):
    """Neighbourhood sums of a random integer layer computed cell by cell through the reference's
    Cell.get_neighborhood (cell.py:202-276) for Moore / von Neumann grids, torus and clamped."""
    import random

    from mesa.discrete_space import OrthogonalMooreGrid, OrthogonalVonNeumannGrid

    rng = np.random.default_rng(0)
    out, cases = {}, []
    k = 0
    for moore in (True, False):
        for torus in (False, True):
            for (rows, cols) in [(7, 9), (5, 5), (12, 8)]:
                for radius in (1, 2, 3):
                    if torus and min(rows, cols) < 2 * radius + 1:
                        continue
                    for include_center in (False, True):
                        klass = OrthogonalMooreGrid if moore else OrthogonalVonNeumannGrid
                        grid = klass((rows, cols), torus=torus, random=random.Random(1))
                        layer = rng.integers(0, 50, size=(rows, cols)).astype(np.int32)
                        sums = np.zeros((rows, cols), dtype=np.int64)
                        sizes = np.zeros((rows, cols), dtype=np.int64)
                        for cell in grid.all_cells:
                            nb = cell.get_neighborhood(radius=radius, include_center=include_center)
                            sums[cell.coordinate] = sum(int(layer[c.coordinate]) for c in nb)
                            sizes[cell.coordinate] = len(nb)
                        out[f"layer_{k}"], out[f"sums_{k}"], out[f"sizes_{k}"] = layer, sums, sizes
                        cases.append((int(moore), int(torus), rows, cols, radius, int(include_center)))
                        k += 1
    out["cases"] = np.array(cases, dtype=np.int64)
    np.savez_compressed(os.path.join(GOLDEN, name), **out)
    print(name, k, "cases")


def neighborhood_nd_fixture(name):
    """The same for 1-, 3- and 4-dimensional grids (grid.py:457-462, 488-501), plus the ordered neighbour lists of a
    few probe cells (connection / BFS order)."""
    import random

    from mesa.discrete_space import OrthogonalMooreGrid, OrthogonalVonNeumannGrid

    rng = np.random.default_rng(1)
    out, cases = {}, []
    k = 0
    for moore in (True, False):
        for torus in (False, True):
            for dims in [(9,), (4, 5, 6), (7, 3, 5), (3, 4, 3, 5)]:
                for radius in (1, 2):
                    if torus and min(dims) < 2 * radius + 1:
                        continue
                    for include_center in (False, True):
                        klass = OrthogonalMooreGrid if moore else OrthogonalVonNeumannGrid
                        grid = klass(dims, torus=torus, random=random.Random(1))
                        layer = rng.integers(0, 50, size=dims).astype(np.int32)
                        sums = np.zeros(dims, dtype=np.int64)
                        sizes = np.zeros(dims, dtype=np.int64)
                        for cell in grid.all_cells:
                            nb = cell.get_neighborhood(radius=radius, include_center=include_center)
                            sums[cell.coordinate] = sum(int(layer[c.coordinate]) for c in nb)
                            sizes[cell.coordinate] = len(nb)
                        probes = [tuple(0 for _ in dims), tuple(d - 1 for d in dims), tuple(d // 2 for d in dims)]
                        for q, pc in enumerate(probes):
                            nb = grid[pc].get_neighborhood(radius=radius, include_center=include_center)
                            out[f"probe_{k}_{q}"] = np.array([c.coordinate for c in nb], dtype=np.int64).reshape(-1, len(dims))
                        out[f"layer_{k}"], out[f"sums_{k}"], out[f"sizes_{k}"] = layer, sums, sizes
                        out[f"dims_{k}"] = np.array(dims, dtype=np.int64)
                        cases.append((int(moore), int(torus), radius, int(include_center)))
                        k += 1
    out["cases"] = np.array(cases, dtype=np.int64)
    np.savez_compressed(os.path.join(GOLDEN, name), **out)
    print(name, k, "cases")


# ------------------------------------------------------------------------------ Boltzmann
def boltzmann_fixture(name, n, width, height, steps, mt_seed, philox_seed):
    from mesa.examples.basic.boltzmann_wealth_model.agents import MoneyAgent
    from mesa.examples.basic.boltzmann_wealth_model.model import BoltzmannScenario, BoltzmannWealth

    with patched_model_random(philox_seed):
        m = BoltzmannWealth(BoltzmannScenario(n=n, width=width, height=height, rng=mt_seed))
    rng = m.random
    agents = sorted(m.agents, key=lambda a: a.unique_id)
    assert [a.unique_id for a in agents] == list(range(1, n + 1))

    def cells():
        return np.array([a.cell.coordinate[0] * height + a.cell.coordinate[1] for a in agents], dtype=np.uint32)

    def wealth():
        return np.array([a.wealth for a in agents], dtype=np.int32)

    out = {"n": n, "width": width, "height": height, "philox_seed": np.uint64(philox_seed), "steps": steps,
           "cell_0": cells(), "wealth_0": wealth()}
    rng.replay = True
    with activation_context(MoneyAgent, "step"):
        for s in range(1, steps + 1):
            m.step()
            out[f"cell_{s}"] = cells()
            out[f"wealth_{s}"] = wealth()
            assert wealth().sum() == n
    rng.replay = False
    # per-cell list order after the last step (arrival order)
    order = np.full(n, -1, dtype=np.int32)
    for cell in m.grid.all_cells:
        for pos, a in enumerate(cell._agents):
            order[a.unique_id - 1] = pos
    out["list_pos_last"] = order
    np.savez_compressed(os.path.join(GOLDEN, name), **out)
    w = wealth()
    print(name, "max wealth", w.max(), "zero-wealth agents", int((w == 0).sum()))


# ------------------------------------------------------------------------------ Boids
def boids_fixture(name, n, width, height, vision, separation, steps, seed, philox_seed):
    """Synchronous goldens: the UNMODIFIED Boid.step of every agent driven on a frozen snapshot
    (call, record, restore).  Inputs are fp32-representable so the device's fp32 state is exact.
    Also a sequential golden: model.step() with the replayed activation order."""
    from mesa.examples.basic.boid_flockers.model import BoidFlockers, BoidsScenario

    def make():
        with patched_model_random(philox_seed):
            m = BoidFlockers(BoidsScenario(population_size=n, width=width, height=height, vision=vision,
                                           separation=separation, rng=seed))
        agents = sorted(m.agents, key=lambda a: a.unique_id)
        assert [a._mesa_index for a in agents] == list(range(n))
        m.space._agent_positions[:n] = m.space._agent_positions[:n].astype(np.float32)
        for a in agents:
            a.direction = np.asarray(a.direction, dtype=np.float64).astype(np.float32).astype(np.float64)
        return m, agents

    m, agents = make()
    out = {"n": n, "size": np.array([width, height], dtype=np.float64), "vision": vision,
           "separation": separation, "cohere": 0.03, "separate": 0.015, "match": 0.05, "speed": 1.0,
           "philox_seed": np.uint64(philox_seed), "steps": steps}
    for t in range(steps):
        pos0 = m.space.agent_positions.copy()
        dir0 = np.array([a.direction for a in agents])
        assert np.array_equal(pos0, pos0.astype(np.float32)) and np.array_equal(dir0, dir0.astype(np.float32))
        pos1, dir1 = np.empty_like(pos0), np.empty_like(dir0)
        counts, idx = [], []
        for i, a in enumerate(agents):
            nb, _ = a.get_neighbors_in_radius(radius=vision)
            counts.append(len(nb))
            idx.extend(b._mesa_index for b in nb)
            a.step()
            pos1[i] = a.position
            dir1[i] = a.direction
            a.position = pos0[i]          # restore the snapshot
            a.direction = dir0[i].copy()
        out[f"pos_in_{t}"] = pos0.astype(np.float32)
        out[f"dir_in_{t}"] = dir0.astype(np.float32)
        out[f"pos_out_{t}"] = pos1
        out[f"dir_out_{t}"] = dir1
        out[f"nbr_count_{t}"] = np.array(counts, dtype=np.int32)
        out[f"nbr_idx_{t}"] = np.array(idx, dtype=np.int32)
        # next snapshot = what an fp32 state would hold
        m.space._agent_positions[:n] = pos1.astype(np.float32)
        for i, a in enumerate(agents):
            a.direction = dir1[i].astype(np.float32).astype(np.float64)

    # sequential (the reference's own shuffle_do semantics) with the replayed activation order
    m, agents = make()
    m.random.replay = True
    out["seq_pos_0"] = m.space.agent_positions.copy()
    out["seq_dir_0"] = np.array([a.direction for a in agents])
    for t in range(1, 4):
        m.step()
        out[f"seq_pos_{t}"] = m.space.agent_positions.copy()
        out[f"seq_dir_{t}"] = np.array([a.direction for a in agents])
    np.savez_compressed(os.path.join(GOLDEN, name), **out)
    print(name, "mean neighbours", float(np.mean(out["nbr_count_0"])))


# ------------------------------------------------------------------------------ continuous-space queries
def continuous_queries_fixture(name):
    """calculate_distances / calculate_difference_vector / get_agents_in_radius / get_k_nearest_agents of the
    UNMODIFIED ContinuousSpace on float32-representable positions (torus and bounded, origin and shifted boxes)."""
    from mesa import Model
    from mesa.experimental.continuous_space import ContinuousSpace, ContinuousSpaceAgent

    out = {}
    cases = []
    rng = np.random.default_rng(2025)
    for c, (n, dims, torus, radius, k) in enumerate([
        (200, [[0, 50], [0, 30]], True, 4.0, 5),
        (200, [[0, 50], [0, 30]], False, 4.0, 5),
        (150, [[-10, 10], [5, 45]], True, 3.0, 8),
        (150, [[-10, 10], [5, 45]], False, 6.5, 1),
        (7, [[0, 1], [0, 1]], True, 0.3, 7),
        (60, [[0, 8], [0, 8]], True, 3.9, 60),
    ]):
        dims = np.asarray(dims, dtype=float)
        model = Model(rng=1)
        space = ContinuousSpace(dims, torus=torus, random=model.random, n_agents=n)
        pos = (dims[:, 0] + rng.random((n, 2)) * (dims[:, 1] - dims[:, 0])).astype(np.float32).astype(np.float64)
        agents = []
        for i in range(n):
            a = ContinuousSpaceAgent(space, model)
            a.position = pos[i]
            agents.append(a)
        index = {id(a): i for i, a in enumerate(agents)}
        points = (dims[:, 0] + rng.random((12, 2)) * (dims[:, 1] - dims[:, 0])).astype(np.float32).astype(np.float64)
        points[0] = pos[3]          # a query sitting on an agent
        out[f"pos_{c}"] = pos
        out[f"points_{c}"] = points
        some = np.sort(rng.choice(n, size=min(n, 9), replace=False))
        out[f"subset_{c}"] = some
        dist_all, delta_all, dist_sub, delta_sub, rad_idx, rad_dist, rad_off, knn_idx, knn_dist = [], [], [], [], [], [], [0], [], []
        for pt in points:
            d, ag = space.calculate_distances(pt)
            assert [index[id(a)] for a in ag] == list(range(n))
            dist_all.append(d)
            delta_all.append(space.calculate_difference_vector(pt))
            sub_agents = [agents[i] for i in some]
            d, _ = space.calculate_distances(pt, sub_agents)
            dist_sub.append(d)
            delta_sub.append(space.calculate_difference_vector(pt, sub_agents))
            ag, d = space.get_agents_in_radius(pt, radius)
            rad_idx.extend(index[id(a)] for a in ag)
            rad_dist.extend(d)
            rad_off.append(len(rad_idx))
            ag, d = space.get_k_nearest_agents(pt, k)
            ids = np.array([index[id(a)] for a in ag])
            o = np.lexsort((ids, d))            # argpartition leaves the k unordered
            knn_idx.append(ids[o])
            knn_dist.append(np.asarray(d)[o])
        out[f"dist_all_{c}"] = np.stack(dist_all)
        out[f"delta_all_{c}"] = np.stack(delta_all)
        out[f"dist_sub_{c}"] = np.stack(dist_sub)
        out[f"delta_sub_{c}"] = np.stack(delta_sub)
        out[f"rad_idx_{c}"] = np.array(rad_idx, dtype=np.int64)
        out[f"rad_dist_{c}"] = np.array(rad_dist, dtype=np.float64)
        out[f"rad_off_{c}"] = np.array(rad_off, dtype=np.int64)
        out[f"knn_idx_{c}"] = np.stack(knn_idx)
        out[f"knn_dist_{c}"] = np.stack(knn_dist)
        # every agent's own neighbours: get_nearest_neighbors(k') excludes self
        kk = min(3, n - 1)
        nn_idx, nn_dist = [], []
        for a in agents[:20]:
            ag, d = a.get_nearest_neighbors(kk)
            ids = np.array([index[id(b)] for b in ag])
            o = np.lexsort((ids, d))
            nn_idx.append(ids[o])
            nn_dist.append(np.asarray(d)[o])
        out[f"nn_idx_{c}"] = np.stack(nn_idx)
        out[f"nn_dist_{c}"] = np.stack(nn_dist)
        cases.append((n, int(torus), radius, k, kk))
        out[f"dims_{c}"] = dims
    out["cases"] = np.array(cases, dtype=np.float64)
    np.savez_compressed(os.path.join(GOLDEN, name), **out)
    print(name, len(cases), "cases")


# ------------------------------------------------------------------------------ cell collections
def cell_collection_fixture(name):
    """CellCollection of a cell's neighbourhood on the live grids (cell_collection.py:33-175): the ordered cells, the
    draws of select_random_cell under a seeded random.Random, and select(filter / at_most)."""
    import random

    from mesa.discrete_space import OrthogonalMooreGrid, OrthogonalVonNeumannGrid

    out, cases = {}, []
    k = 0
    for moore in (True, False):
        for torus in (False, True):
            for radius in (1, 2):
                klass = OrthogonalMooreGrid if moore else OrthogonalVonNeumannGrid
                grid = klass((7, 5), torus=torus, random=random.Random(123))
                for probe in [(0, 0), (3, 2), (6, 4)]:
                    nb = grid[probe].get_neighborhood(radius=radius)
                    out[f"cells_{k}"] = np.array([c.coordinate for c in nb.cells], dtype=np.int64)
                    out[f"draws_{k}"] = np.array([nb.select_random_cell().coordinate for _ in range(6)], dtype=np.int64)
                    even = nb.select(lambda c: (c.coordinate[0] + c.coordinate[1]) % 2 == 0)
                    out[f"even_{k}"] = np.array([c.coordinate for c in even.cells], dtype=np.int64).reshape(-1, 2)
                    out[f"first3_{k}"] = np.array([c.coordinate for c in nb.select(at_most=3).cells], dtype=np.int64)
                    out[f"half_{k}"] = np.array([c.coordinate for c in nb.select(at_most=0.5).cells], dtype=np.int64).reshape(-1, 2)
                    cases.append((int(moore), int(torus), radius, probe[0], probe[1]))
                    k += 1
    out["cases"] = np.array(cases, dtype=np.int64)
    np.savez_compressed(os.path.join(GOLDEN, name), **out)
    print(name, k, "cases")


# ------------------------------------------------------------------------------ n-d continuous spaces
def continuous_nd_fixture(name):
    """The point queries of the UNMODIFIED ContinuousSpace in 1, 3 and 4 dimensions (torus and bounded, shifted boxes)
    on float32-representable positions: all distances, difference vectors, radius selections, k nearest."""
    from mesa import Model
    from mesa.experimental.continuous_space import ContinuousSpace, ContinuousSpaceAgent

    out = {}
    cases = []
    rng = np.random.default_rng(77)
    for c, (n, dims, torus, radius, k) in enumerate([
        (120, [[0, 20], [0, 30], [0, 10]], True, 5.0, 6),
        (120, [[0, 20], [0, 30], [0, 10]], False, 5.0, 6),
        (90, [[-5, 5], [2, 12], [-30, -10], [0, 7]], True, 6.0, 4),
        (90, [[-5, 5], [2, 12], [-30, -10], [0, 7]], False, 7.5, 90),
        (40, [[3, 19]], True, 1.5, 3),
        (40, [[3, 19]], False, 1.5, 40),
        (5, [[0, 2], [-2, 0], [-2, 2]], False, 1.0, 2),
    ]):
        dims = np.asarray(dims, dtype=float)
        nd = dims.shape[0]
        model = Model(rng=1)
        space = ContinuousSpace(dims, torus=torus, random=model.random, n_agents=n)
        pos = (dims[:, 0] + rng.random((n, nd)) * (dims[:, 1] - dims[:, 0])).astype(np.float32).astype(np.float64)
        if n > 10:
            pos[7] = pos[2]         # two agents on the same spot: ties rank by insertion order
        agents = []
        for i in range(n):
            a = ContinuousSpaceAgent(space, model)
            a.position = pos[i]
            agents.append(a)
        index = {id(a): i for i, a in enumerate(agents)}
        points = (dims[:, 0] + rng.random((8, nd)) * (dims[:, 1] - dims[:, 0])).astype(np.float32).astype(np.float64)
        points[0] = pos[3 % n]
        some = np.sort(rng.choice(n, size=min(n, 7), replace=False))
        dist_all, delta_all, dist_sub, delta_sub, rad_idx, rad_dist, rad_off, knn_idx, knn_dist = [], [], [], [], [], [], [0], [], []
        for pt in points:
            d, ag = space.calculate_distances(pt)
            assert [index[id(a)] for a in ag] == list(range(n))
            dist_all.append(d)
            delta_all.append(space.calculate_difference_vector(pt))
            sub_agents = [agents[i] for i in some]
            d, _ = space.calculate_distances(pt, sub_agents)
            dist_sub.append(d)
            delta_sub.append(space.calculate_difference_vector(pt, sub_agents))
            ag, d = space.get_agents_in_radius(pt, radius)
            rad_idx.extend(index[id(a)] for a in ag)
            rad_dist.extend(d)
            rad_off.append(len(rad_idx))
            ag, d = space.get_k_nearest_agents(pt, k)
            ids = np.array([index[id(a)] for a in ag])
            o = np.lexsort((ids, d))            # argpartition leaves the k unordered
            knn_idx.append(ids[o])
            knn_dist.append(np.asarray(d)[o])
        out[f"pos_{c}"], out[f"points_{c}"], out[f"subset_{c}"], out[f"dims_{c}"] = pos, points, some, dims
        out[f"dist_all_{c}"] = np.stack(dist_all)
        out[f"delta_all_{c}"] = np.stack(delta_all)
        out[f"dist_sub_{c}"] = np.stack(dist_sub)
        out[f"delta_sub_{c}"] = np.stack(delta_sub)
        out[f"rad_idx_{c}"] = np.array(rad_idx, dtype=np.int64)
        out[f"rad_dist_{c}"] = np.array(rad_dist, dtype=np.float64)
        out[f"rad_off_{c}"] = np.array(rad_off, dtype=np.int64)
        out[f"knn_idx_{c}"] = np.stack(knn_idx)
        out[f"knn_dist_{c}"] = np.stack(knn_dist)
        cases.append((n, int(torus), radius, k))
    out["cases"] = np.array(cases, dtype=np.float64)
    np.savez_compressed(os.path.join(GOLDEN, name), **out)
    print(name, len(cases), "cases")


# ------------------------------------------------------------------------------ dynamic populations
def dynamic_agents_fixture(name):
    """A scripted sequence of agent creations (Agent.create_agents, agent.py:115-160) and removals (Agent.remove,
    agent.py:73-108) on the live Model; after every operation the (unique_id, x) columns of model.agents in
    iteration order -- what DeviceAgentSet.add_agents / remove_agents must reproduce."""
    import mesa

    class Walker(mesa.Agent):
        def __init__(self, model, x=0):
            super().__init__(model)
            self.x = x

    m = mesa.Model(rng=1)
    script = [("add", np.arange(40) * 3), ("remove_x_mod4", None), ("add", 1000 + np.arange(7)),
              ("remove_ids", np.array([2, 41, 47, 39])), ("add", np.array([5, 6, 7])), ("remove_ids", np.array([1])),
              ("remove_gt", np.array([45])), ("add", np.array([77])), ("remove_ids", np.array([], dtype=np.int64))]
    out = {"n_ops": np.array(len(script))}
    for k, (kind, arg) in enumerate(script):
        if kind == "add":
            Walker.create_agents(m, len(arg), x=list(arg))
            removed = np.array([], dtype=np.int64)
        else:
            if kind == "remove_x_mod4":
                victims = [a for a in m.agents if a.x % 4 == 0]
            elif kind == "remove_gt":
                victims = [a for a in m.agents if a.unique_id > int(arg[0])]
            else:
                wanted = set(int(i) for i in arg)
                victims = [a for a in m.agents if a.unique_id in wanted]
            removed = np.array([a.unique_id for a in victims], dtype=np.int64)
            for a in victims:
                a.remove()
        out[f"op{k}_kind"] = np.array(0 if kind == "add" else 1)
        out[f"op{k}_add_x"] = np.asarray(arg if kind == "add" else [], dtype=np.int64)
        out[f"op{k}_removed_ids"] = removed
        out[f"op{k}_ids"] = np.array(m.agents.get("unique_id"), dtype=np.int64)
        out[f"op{k}_x"] = np.array(m.agents.get("x"), dtype=np.int64)
        assert list(out[f"op{k}_ids"]) == [a.unique_id for a in m.agents_by_type[Walker]]
    np.savez_compressed(os.path.join(GOLDEN, name), **out)
    print(name, "final population", out[f"op{len(script) - 1}_ids"].tolist())


# ------------------------------------------------------------------------------ Wolf-Sheep
def wolf_sheep_fixture(name, width, height, n_sheep, n_wolves, steps, mt_seed, philox_seed, grass=True, regrowth=30,
                       sheep_reproduce=0.04, wolf_reproduce=0.05, wolf_gain=20.0):
    """The unmodified WolfSheep model (examples/advanced/wolf_sheep) under the replayed order / draws: after every step
    the sheep and wolf tables (unique_id, cell, energy) in registration order and the grass layer."""
    from mesa.examples.advanced.wolf_sheep.agents import Animal, GrassPatch, Sheep, Wolf
    from mesa.examples.advanced.wolf_sheep.model import WolfSheep, WolfSheepScenario

    with patched_model_random(philox_seed):
        m = WolfSheep(WolfSheepScenario(width=width, height=height, initial_sheep=n_sheep, initial_wolves=n_wolves,
                                        grass=grass, grass_regrowth_time=regrowth, sheep_reproduce=sheep_reproduce,
                                        wolf_reproduce=wolf_reproduce, wolf_gain_from_food=float(wolf_gain),
                                        rng=mt_seed))
    rng = m.random
    assert isinstance(rng, ReplayRandom)
    H, W = m.grid.dimensions          # model.py:76: Grid([height, width])

    def table(klass):
        agents = list(m.agents_by_type.get(klass, []))
        ids = np.array([a.unique_id for a in agents], dtype=np.int64)
        cells = np.array([a.cell.coordinate[0] * W + a.cell.coordinate[1] for a in agents], dtype=np.int64)
        energy = np.array([a.energy for a in agents], dtype=np.float64)
        return ids, cells, energy

    def grown():
        g = np.zeros((H, W), dtype=np.uint8)
        if grass:
            for a in m.agents_by_type[GrassPatch]:
                g[a.cell.coordinate] = a.fully_grown
        return g

    out = {"height": H, "width": W, "grass": grass, "regrowth": regrowth, "steps": steps, "mt_seed": mt_seed,
           "philox_seed": np.uint64(philox_seed), "sheep_reproduce": sheep_reproduce, "wolf_reproduce": wolf_reproduce,
           "sheep_gain": m.scenario.sheep_gain_from_food, "wolf_gain": m.scenario.wolf_gain_from_food,
           "next_id_0": m.agent_id_counter}
    for kind, klass in (("sheep", Sheep), ("wolf", Wolf)):
        out[f"{kind}_id_0"], out[f"{kind}_cell_0"], out[f"{kind}_energy_0"] = table(klass)
    out["grown_0"] = grown()
    # the pending regrowth events of the initial patches (agents.py:142-144): time at which each patch regrows, -1 if grown
    regrow_at = np.full((H, W), -1, dtype=np.int64)
    for ev in m._event_list._events if hasattr(m._event_list, "_events") else []:
        fn = getattr(ev, "fn", None)
        target = fn() if callable(fn) else None
        owner = getattr(target, "__self__", None)
        if isinstance(owner, GrassPatch) and not getattr(ev, "CANCELED", False):
            regrow_at[owner.cell.coordinate] = int(ev.time)
    out["regrow_at_0"] = regrow_at
    rng.replay = True
    with activation_context(Animal, "step"):
        for s in range(1, steps + 1):
            m.step()
            for kind, klass in (("sheep", Sheep), ("wolf", Wolf)):
                out[f"{kind}_id_{s}"], out[f"{kind}_cell_{s}"], out[f"{kind}_energy_{s}"] = table(klass)
            out[f"grown_{s}"] = grown()
    rng.replay = False
    out["counts"] = m.datacollector.get_model_vars_dataframe().to_numpy()
    np.savez_compressed(os.path.join(GOLDEN, name), **out)
    print(name, "sheep", [len(out[f"sheep_id_{s}"]) for s in (0, steps // 2, steps)], "wolves",
          [len(out[f"wolf_id_{s}"]) for s in (0, steps // 2, steps)], "pending regrowths", int((regrow_at >= 0).sum()))


def wolf_sheep_fixtures():
    wolf_sheep_fixture("wolf_sheep_20x20.npz", 20, 20, 100, 50, 25, 42, 17)                      # the scenario's defaults
    wolf_sheep_fixture("wolf_sheep_30x30.npz", 30, 30, 300, 25, 40, 1, 21, regrowth=20, sheep_reproduce=0.1,
                       wolf_gain=15.0)                                                           # both species persist
    wolf_sheep_fixture("wolf_sheep_12x15_crowded.npz", 15, 12, 400, 20, 20, 3, 18, regrowth=8, sheep_reproduce=0.2,
                       wolf_gain=10.0)                                                           # crowded cells, many births
    wolf_sheep_fixture("wolf_sheep_9x7_nograss.npz", 7, 9, 60, 6, 15, 5, 19, grass=False, sheep_reproduce=0.3)


def epstein_fixture(name, width, height, steps, mt_seed, philox_seed, **params):
    """The unmodified EpsteinCivilViolence model (examples/advanced/epstein_civil_violence) under the replayed order / draws:
    after every step every agent's cell, state, jail sentence and arrest probability (rows = unique_id - 1)."""
    from mesa.examples.advanced.epstein_civil_violence.agents import Citizen, Cop
    from mesa.examples.advanced.epstein_civil_violence.model import EpsteinCivilViolence, EpsteinScenario

    with patched_model_random(philox_seed):
        m = EpsteinCivilViolence(width=width, height=height, scenario=EpsteinScenario(rng=mt_seed, **params))
    rng = m.random
    assert isinstance(rng, ReplayRandom)
    W, H = m.grid.dimensions
    agents = sorted(m.agents, key=lambda a: a.unique_id)
    assert [a.unique_id for a in agents] == list(range(1, len(agents) + 1))

    def table():
        cell = np.array([a.cell.coordinate[0] * H + a.cell.coordinate[1] for a in agents], dtype=np.int64)
        state = np.array([a.state.value if isinstance(a, Citizen) else 0 for a in agents], dtype=np.int8)
        jail = np.array([a.jail_sentence if isinstance(a, Citizen) else 0 for a in agents], dtype=np.int64)
        ap = np.array([np.nan if getattr(a, "arrest_probability", None) is None else a.arrest_probability for a in agents],
                      dtype=np.float64)
        return cell, state, jail, ap

    sc = m.scenario
    out = {"width": W, "height": H, "steps": steps, "mt_seed": mt_seed, "philox_seed": np.uint64(philox_seed),
           "moore": sc.grid_type == "Moore", "vision": sc.cop_vision, "legitimacy": sc.legitimacy,
           "max_jail_term": sc.max_jail_term, "active_threshold": sc.active_threshold,
           "arrest_prob_constant": sc.arrest_prob_constant, "movement": sc.movement,
           "citizen_density": sc.citizen_density, "cop_density": sc.cop_density,
           "kind": np.array([isinstance(a, Cop) for a in agents], dtype=np.int8),
           "hardship": np.array([getattr(a, "hardship", 0.0) for a in agents], dtype=np.float64),
           "risk_aversion": np.array([getattr(a, "risk_aversion", 0.0) for a in agents], dtype=np.float64)}
    out["cell_0"], out["state_0"], out["jail_0"], _ = table()
    rng.replay = True
    with activation_context(Citizen, "step"), activation_context(Cop, "step"):
        for s in range(1, steps + 1):
            m.step()
            out[f"cell_{s}"], out[f"state_{s}"], out[f"jail_{s}"], out[f"ap_{s}"] = table()
    rng.replay = False
    out["counts"] = m.datacollector.get_model_vars_dataframe().to_numpy()
    np.savez_compressed(os.path.join(GOLDEN, name), **out)
    print(name, len(agents), "agents; active / quiet / arrested:", [tuple(r) for r in out["counts"][[0, 1, steps // 2, steps]]])


def epstein_fixtures():
    epstein_fixture("epstein_40x40.npz", 40, 40, 12, 42, 31)                                  # the scenario's defaults (von Neumann, vision 7)
    epstein_fixture("epstein_30x22_lively.npz", 30, 22, 25, 3, 32, legitimacy=0.55, max_jail_term=6, cop_density=0.04,
                    cop_vision=4, citizen_vision=4)                                            # rebellions flare up and are put down
    epstein_fixture("epstein_17x19_moore.npz", 17, 19, 15, 5, 33, grid_type="Moore", cop_vision=3, citizen_vision=3,
                    legitimacy=0.6, max_jail_term=3, citizen_density=0.8, cop_density=0.1)     # Moore, crowded: few empty cells
    epstein_fixture("epstein_16x16_still.npz", 16, 16, 8, 7, 34, movement=False, legitimacy=0.5, max_jail_term=4,
                    cop_vision=2, citizen_vision=2)                                            # nobody moves


def sugarscape_fixture(name, steps, mt_seed, philox_seed, trade=False, **params):
    """The unmodified SugarscapeG1mt model (examples/advanced/sugarscape_g1mt) with enable_trade=False under the replayed
    order / draws: after every step the traders (unique ids in model.agents order, cells, sugar, spice) and both layers."""
    from mesa.examples.advanced.sugarscape_g1mt.agents import Trader
    from mesa.examples.advanced.sugarscape_g1mt.model import SugarscapeG1mt, SugarScapeScenario

    with patched_model_random(philox_seed):
        m = SugarscapeG1mt(SugarScapeScenario(rng=mt_seed, enable_trade=trade, **params))
    rng = m.random
    assert isinstance(rng, ReplayRandom)
    W, H = m.grid.dimensions

    def table():
        agents = list(m.agents)
        return (np.array([a.unique_id for a in agents], dtype=np.int64),
                np.array([a.cell.coordinate[0] * H + a.cell.coordinate[1] for a in agents], dtype=np.int64),
                np.array([a.sugar for a in agents], dtype=np.float64), np.array([a.spice for a in agents], dtype=np.float64))

    agents = list(m.agents)
    out = {"width": W, "height": H, "steps": steps, "mt_seed": mt_seed, "philox_seed": np.uint64(philox_seed),
           "sugar_max": np.asarray(m.sugar_distribution, dtype=np.float64), "spice_max": np.asarray(m.spice_distribution, dtype=np.float64),
           "metabolism_sugar": np.array([a.metabolism_sugar for a in agents], dtype=np.int64),
           "metabolism_spice": np.array([a.metabolism_spice for a in agents], dtype=np.int64),
           "vision": np.array([a.vision for a in agents], dtype=np.int64),
           "scenario": np.array([m.scenario.initial_population, m.scenario.endowment_min, m.scenario.endowment_max,
                                 m.scenario.metabolism_min, m.scenario.metabolism_max, m.scenario.vision_min, m.scenario.vision_max])}
    out["id_0"], out["cell_0"], out["sugar_0"], out["spice_0"] = table()
    rng.replay = True
    with activation_context(Trader, "step"), activation_context(Trader, "trade_with_neighbors"):
        for s in range(1, steps + 1):
            m.step()
            out[f"id_{s}"], out[f"cell_{s}"], out[f"sugar_{s}"], out[f"spice_{s}"] = table()
            if trade:   # this step's trades, flattened in model.agents order: (trader id, partner id, price)
                rec = [(a.unique_id, p, pr) for a in m.agents for p, pr in zip(a.trade_partners, a.prices)]
                out[f"trades_{s}"] = np.array(rec, dtype=np.float64).reshape(-1, 3)
            if s in (1, steps // 2, steps):
                out[f"sugar_layer_{s}"] = np.asarray(m.grid.sugar.data if hasattr(m.grid.sugar, "data") else m.grid.sugar, dtype=np.float64).copy()
                out[f"spice_layer_{s}"] = np.asarray(m.grid.spice.data if hasattr(m.grid.spice, "data") else m.grid.spice, dtype=np.float64).copy()
    rng.replay = False
    out["counts"] = m.datacollector.get_model_vars_dataframe().to_numpy()
    np.savez_compressed(os.path.join(GOLDEN, name), **out)
    print(name, "traders", [len(out[f"id_{s}"]) for s in (0, 1, steps // 2, steps)])


def sugarscape_fixtures():
    sugarscape_fixture("sugarscape_200.npz", 30, 42, 41)                                                  # the scenario's defaults
    sugarscape_fixture("sugarscape_600_crowded.npz", 20, 7, 43, initial_population=600, vision_max=3)     # shared cells at the start
    sugarscape_fixture("sugarscape_80_farsighted.npz", 25, 3, 44, initial_population=80, vision_min=4, vision_max=5,
                       metabolism_max=3)
    sugarscape_fixture("sugarscape_200_trade.npz", 30, 42, 45, trade=True)                                # the scenario's defaults
    sugarscape_fixture("sugarscape_500_trade_crowded.npz", 15, 11, 46, trade=True, initial_population=500, vision_max=4)


def main():
    import_reference()
    os.makedirs(GOLDEN, exist_ok=True)
    only = sys.argv[1] if len(sys.argv) > 1 else ""
    if only == "neighborhoods":
        neighborhood_fixture("neighborhoods.npz")
        return
    if only == "neighborhoods_nd":
        neighborhood_nd_fixture("neighborhoods_nd.npz")
        return
    if only == "continuous":
        continuous_queries_fixture("continuous_queries.npz")
        return
    if only == "cell_collection":
        cell_collection_fixture("cell_collection.npz")
        return
    if only == "continuous_nd":
        continuous_nd_fixture("continuous_queries_nd.npz")
        return
    if only == "dynamic":
        dynamic_agents_fixture("dynamic_agents.npz")
        return
    if only == "wolf_sheep":
        wolf_sheep_fixtures()
        return
    if only == "epstein":
        epstein_fixtures()
        return
    if only == "sugarscape":
        sugarscape_fixtures()
        return
    if only == "boltzmann":
        boltzmann_fixture("boltzmann_100_10x10.npz", 100, 10, 10, 25, 42, 3)      # reference "small"
        boltzmann_fixture("boltzmann_2000_20x30.npz", 2000, 20, 30, 10, 7, 4)     # crowded cells
        boltzmann_fixture("boltzmann_10000_100x100.npz", 10000, 100, 100, 10, 1, 5)  # reference "large"
        return
    if only == "datacollector":
        datacollector_fixture("datacollector_schelling_20x20.npz", 20, 20, 0.8, 0.4, 1, 10, 42, 2024)
        return
    if only == "boids":
        boids_fixture("boids_300.npz", 300, 100, 100, 10.0, 2.0, 3, 42, 5)
        boids_fixture("boids_120_wrap.npz", 120, 40, 30, 6.0, 2.0, 3, 1, 6)
        return
    neighborhood_fixture("neighborhoods.npz")
    neighborhood_nd_fixture("neighborhoods_nd.npz")
    gol_fixture(37, 23, 42, 6, "gol_37x23.npz")
    gol_fixture(64, 48, 7, 20, "gol_64x48.npz")
    schelling_fixture("schelling_20x20_r1.npz", 20, 20, 0.8, 0.4, 1, 30, 42, 2024, full_steps=range(1, 31))
    schelling_fixture("schelling_16x48_r2.npz", 16, 48, 0.9, 0.6, 2, 12, 5, 99, full_steps=range(1, 13))
    datacollector_fixture("datacollector_schelling_20x20.npz", 20, 20, 0.8, 0.4, 1, 10, 42, 2024)
    boids_fixture("boids_300.npz", 300, 100, 100, 10.0, 2.0, 3, 42, 5)
    boids_fixture("boids_120_wrap.npz", 120, 40, 30, 6.0, 2.0, 3, 1, 6)
    continuous_queries_fixture("continuous_queries.npz")
    continuous_nd_fixture("continuous_queries_nd.npz")
    cell_collection_fixture("cell_collection.npz")
    dynamic_agents_fixture("dynamic_agents.npz")
    boltzmann_fixture("boltzmann_100_10x10.npz", 100, 10, 10, 25, 42, 3)
    boltzmann_fixture("boltzmann_2000_20x30.npz", 2000, 20, 30, 10, 7, 4)
    boltzmann_fixture("boltzmann_10000_100x100.npz", 10000, 100, 100, 10, 1, 5)
    wolf_sheep_fixtures()
    epstein_fixtures()
    sugarscape_fixtures()
    # BASELINE.json configs[0]: 100x100, density 0.8, 100 steps (h=0.4, r=1)
    schelling_fixture("schelling_100x100_c1.npz", 100, 100, 0.8, 0.4, 1, 100, 42, 7, full_steps=[1, 2, 10, 50, 100])
    # the reference's own "large" benchmark variant (benchmarks/configurations.py:42-49)
    schelling_fixture("schelling_100x100_large.npz", 100, 100, 0.8, 1, 2, 10, 42, 11, full_steps=[1, 5, 10])


if __name__ == "__main__":
    main()
