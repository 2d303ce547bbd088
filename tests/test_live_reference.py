"""Live differential runs against the UNMODIFIED reference on seeds that are in no fixture (build container only: skipped
where /root/reference is absent, e.g. on the GPU box -- nothing here is marked gpu).

From nothing but a seed: the reference model is constructed from `rng=seed` and stepped with the replayed Philox
activation order / draws (oracle/replay.py); on our side the initial state comes from mesa_b200.reference_init (the
reference's seeding and per-cell draw protocol, vectorised) and the trajectory from the CPU oracle.  Equality at every
step pins both the seed-to-initial-state path and the oracle on inputs the golden fixtures do not contain."""
import os
import random

import numpy as np
import pytest

import oracle
from mesa_b200 import reference_init as ri

pytestmark = pytest.mark.skipif(not os.path.isdir("/root/reference/mesa"), reason="needs the reference checkout")


@pytest.fixture(scope="module")
def replay():
    from oracle import replay as rp

    rp.import_reference()
    return rp


@pytest.mark.parametrize("seed,width,height,density,homophily,radius", [(2027, 15, 12, 0.75, 0.5, 1), (31337, 9, 21, 0.9, 0.7, 2),
                                                                          (5150, 12, 12, 0.97, 0.6, 1)])  # dense: 50 probes miss -> argwhere fallback
def test_schelling_from_a_seed(replay, seed, width, height, density, homophily, radius):
    from mesa.examples.basic.schelling.agents import SchellingAgent
    from mesa.examples.basic.schelling.model import Schelling, SchellingScenario

    philox_seed = seed + 1
    with replay.patched_model_random(philox_seed):
        m = Schelling(SchellingScenario(width=width, height=height, density=density, minority_pc=0.5,
                                        homophily=homophily, radius=radius, rng=seed))
    agents = sorted(m.agents, key=lambda a: a.unique_id)

    def ref_cells():
        return np.array([a.cell.coordinate[0] * height + a.cell.coordinate[1] for a in agents], dtype=np.int64)

    types = ri.schelling_initial_types(random.Random(ri.seed_streams(seed)[1]), width, height, density, 0.5)
    st = oracle.SchellingState.from_type_grid(types)
    assert st.n == len(agents) and np.array_equal(st.cell, ref_cells())
    assert np.array_equal(st.atype, np.array([a.type for a in agents], dtype=np.int8))
    assert st.assign_state(homophily, radius) == m.happy
    m.random.replay = True
    with replay.activation_context(SchellingAgent, "step"):
        for step in range(6):
            m.step()
            st.move_unhappy(philox_seed, step)
            assert st.assign_state(homophily, radius) == m.happy, step
            assert np.array_equal(st.cell, ref_cells()), step
            assert np.array_equal(st.happy, np.array([a.happy for a in agents], dtype=np.uint8)), step


@pytest.mark.parametrize("seed,width,height", [(99, 21, 34), (123456, 40, 17)])
def test_game_of_life_from_a_seed(replay, seed, width, height):
    from mesa.examples.basic.conways_game_of_life.model import ConwaysGameOfLife

    m = ConwaysGameOfLife(width=width, height=height, rng=seed)

    def ref_state():
        s = np.zeros((width, height), np.uint8)
        for a in m.agents:
            s[a.cell.coordinate] = a.state
        return s

    s = ri.gol_initial_state(random.Random(ri.seed_streams(seed)[1]), width, height, 0.2).astype(np.uint8)
    assert np.array_equal(s, ref_state())
    for step in range(5):
        m.step()
        s = oracle.gol_step(s, True)
        assert np.array_equal(s, ref_state()), step


@pytest.mark.parametrize("seed,n,width,height", [(77, 300, 12, 9), (4242, 1500, 25, 40)])
def test_boltzmann_from_a_seed(replay, seed, n, width, height):
    from mesa.examples.basic.boltzmann_wealth_model.agents import MoneyAgent
    from mesa.examples.basic.boltzmann_wealth_model.model import BoltzmannScenario, BoltzmannWealth

    philox_seed = seed + 5
    with replay.patched_model_random(philox_seed):
        m = BoltzmannWealth(BoltzmannScenario(n=n, width=width, height=height, rng=seed))
    agents = sorted(m.agents, key=lambda a: a.unique_id)

    def ref_cells():
        return np.array([a.cell.coordinate[0] * height + a.cell.coordinate[1] for a in agents], dtype=np.int64)

    cells = ri.boltzmann_initial_cells(random.Random(ri.seed_streams(seed)[1]), n, width * height)
    assert np.array_equal(cells, ref_cells())
    st = oracle.BoltzmannState(width, height, cells)
    m.random.replay = True
    with replay.activation_context(MoneyAgent, "step"):
        for step in range(5):
            m.step()
            st.step(philox_seed, step)
            assert np.array_equal(st.cell, ref_cells()), step
            assert np.array_equal(st.wealth, np.array([a.wealth for a in agents], dtype=np.int32)), step


def test_boids_initial_state_from_a_seed(replay):
    from mesa.examples.basic.boid_flockers.model import BoidFlockers, BoidsScenario

    m = BoidFlockers(BoidsScenario(population_size=50, width=60, height=40, rng=2028))
    gen, _ = ri.seed_streams(2028)
    pos = gen.random(size=(50, 2)) * np.array([60.0, 40.0])               # boid_flockers/model.py:79-82
    dirs = gen.uniform(-1, 1, size=(50, 2))
    agents = sorted(m.agents, key=lambda a: a.unique_id)
    assert np.array_equal(np.array([a.position for a in agents]), pos)
    assert np.array_equal(np.array([a.direction for a in agents]), dirs)


def test_agentset_operations_against_the_reference_agentset(replay, host_kernels):
    """select / sort / get / groupby / agg of DeviceAgentSet (column predicates on CPU tensors; the device sort and
    reductions replaced by host doubles) against mesa.AgentSet on the same random attribute table."""
    import mesa
    import torch

    from mesa_b200 import ops
    from mesa_b200.agentset import DeviceAgent, DeviceAgentSet
    from mesa_b200.model import DeviceModel


    class RefAgent(mesa.Agent):
        def __init__(self, model, energy, kind):
            super().__init__(model)
            self.energy, self.kind = energy, kind

    class DevAgent(DeviceAgent):
        pass

    rng = np.random.default_rng(5)
    for trial in range(4):
        n = int(rng.integers(5, 60))
        energy = rng.integers(0, 6, size=n)
        kind = rng.integers(0, 4, size=n)
        ref_model = mesa.Model(rng=1)
        RefAgent.create_agents(ref_model, n, list(energy), list(kind))
        ref = ref_model.agents
        dev = DeviceAgentSet(DeviceModel(rng=1), DevAgent, n, {"energy": torch.from_numpy(energy), "kind": torch.from_numpy(kind)})

        def ids(agentset):
            return [a.unique_id for a in agentset]

        for at_most in (3, 0.3, 1.0, 1, float("inf")):
            for thresh in (None, 2):
                rf = None if thresh is None else (lambda a, t=thresh: a.energy >= t)
                df = None if thresh is None else (lambda c, t=thresh: c["energy"] >= t)
                assert ids(dev.select(df, at_most=at_most)) == ids(ref.select(rf, at_most=at_most)), (trial, at_most, thresh)
        for ascending in (True, False):
            assert ids(dev.sort("energy", ascending=ascending)) == ids(ref.sort("energy", ascending=ascending))
        assert dev.get("energy").tolist() == ref.get("energy")
        assert dev.get(["energy", "kind"]).tolist() == ref.get(["energy", "kind"])
        rg, dg = ref.groupby("kind"), dev.groupby("kind")
        assert list(dg.groups) == list(rg.groups) and dg.count() == rg.count()
        for k in rg.groups:
            assert ids(dg.get_group(k)) == ids(rg.get_group(k))
        assert dg.agg("energy", sum) == rg.agg("energy", sum) and dg.agg("energy", max) == rg.agg("energy", max)
        assert dev.agg("energy", [min, max, sum]) == ref.agg("energy", [min, max, sum])
        assert dev.agg("energy", np.mean) == ref.agg("energy", np.mean)


def test_continuous_space_agent_queries_against_the_reference(replay, monkeypatch):
    """ContinuousSpaceAgent.get_neighbors_in_radius / get_nearest_neighbors / position setter
    (continuous_space_agents.py:31-91) through DeviceContinuousSpaceAgent row proxies.  The space runs on CPU tensors
    here with the two device queries replaced by oracle doubles (the kernels themselves are pinned by
    tests/test_continuous_gpu.py); what is compared is the agent-level logic: self dropped, k + 1 asked for, order."""
    import mesa
    import torch
    from mesa.experimental.continuous_space import ContinuousSpace, ContinuousSpaceAgent

    from mesa_b200 import ops
    from mesa_b200.agentset import DeviceAgentSet
    from mesa_b200.continuous_space import DeviceContinuousSpace, DeviceContinuousSpaceAgent
    from mesa_b200.model import DeviceModel

    def host_radius_query(pos, queries, size, radius, torus=True, origin=(0.0, 0.0), exclude=None):
        off, idx, dist = [0], [], []
        p = pos.numpy().astype(np.float64)
        for q in queries.numpy().astype(np.float64):
            d = oracle.point_distances(p, q, size, torus)
            inside = np.nonzero(d <= radius)[0]
            idx.extend(inside.tolist())
            dist.extend(d[inside].tolist())
            off.append(len(idx))
        return torch.tensor(off), torch.tensor(idx, dtype=torch.int64), torch.tensor(dist, dtype=torch.float64)

    def host_knn_query(pos, queries, size, k, torus=True, origin=(0.0, 0.0), exclude=None):
        p = pos.numpy().astype(np.float64)
        rows = [oracle.knn_query(p, q, size, k, torus) for q in queries.numpy().astype(np.float64)]
        return torch.tensor(np.stack([r[0] for r in rows])), torch.tensor(np.stack([r[1] for r in rows]))

    monkeypatch.setattr(ops, "radius_query", host_radius_query)
    monkeypatch.setattr(ops, "knn_query", host_knn_query)

    class Bird(DeviceContinuousSpaceAgent):
        pass

    rng = np.random.default_rng(11)
    for torus in (True, False):
        dims = np.array([[0.0, 30.0], [0.0, 20.0]])
        n = 60
        pos = (rng.random((n, 2)) * (dims[:, 1] - dims[:, 0])).astype(np.float32).astype(np.float64)
        ref_model = mesa.Model(rng=1)
        ref_space = ContinuousSpace(dims, torus=torus, random=ref_model.random, n_agents=n)
        ref_agents = []
        for i in range(n):
            a = ContinuousSpaceAgent(ref_space, ref_model)
            a.position = pos[i]
            ref_agents.append(a)
        index = {id(a): i for i, a in enumerate(ref_agents)}

        model = DeviceModel(rng=1)
        model.space = DeviceContinuousSpace(dims, torus=torus, random=model.random, n_agents=n, device="cpu")
        model.space.add_agents(pos)
        birds = DeviceAgentSet(model, Bird, n, {"position": model.space._agent_positions}, resizable=False)
        for i in range(0, n, 7):
            ag, d = ref_agents[i].get_neighbors_in_radius(6.0)
            got_ag, got_d = birds[i].get_neighbors_in_radius(6.0)
            assert [b.unique_id - 1 for b in got_ag] == [index[id(a)] for a in ag] and np.array_equal(got_d, d)
            ag, d = ref_agents[i].get_nearest_neighbors(4)
            got_ag, got_d = birds[i].get_nearest_neighbors(4)
            o = np.lexsort(([index[id(a)] for a in ag], d))                  # argpartition leaves the k unordered
            assert [b.unique_id - 1 for b in got_ag] == [index[id(ag[j])] for j in o] and np.array_equal(got_d, np.asarray(d)[o])
        # position setter: wraps on a torus, refuses out-of-bounds points otherwise
        if torus:
            ref_agents[0].position = np.array([31.5, -2.0])
            birds[0].position = np.array([31.5, -2.0])
            assert np.allclose(birds[0].position, ref_agents[0].position) and np.allclose(birds[0].position, [1.5, 18.0])
        else:
            with pytest.raises(ValueError):
                ref_agents[0].position = np.array([31.5, 1.0])
            with pytest.raises(ValueError):
                birds[0].position = np.array([31.5, 1.0])


@pytest.mark.parametrize("seed,width,height,params", [
    (77, 24, 19, dict(legitimacy=0.6, max_jail_term=5, cop_vision=4, citizen_vision=4)),
    (4242, 15, 16, dict(grid_type="Moore", legitimacy=0.5, max_jail_term=2, cop_vision=2, citizen_vision=2, cop_density=0.1)),
])
def test_epstein_from_a_seed(replay, seed, width, height, params):
    """Epstein civil violence: initial state from the seed (reference_init.epstein_initial_state), trajectory from the
    sequential restatement (oracle/epstein.py) == the unmodified model under the replayed order / draws."""
    from mesa.examples.advanced.epstein_civil_violence.agents import Citizen, Cop
    from mesa.examples.advanced.epstein_civil_violence.model import EpsteinCivilViolence, EpsteinScenario
    from oracle.epstein import EpsteinOracle

    philox_seed = seed + 1
    with replay.patched_model_random(philox_seed):
        m = EpsteinCivilViolence(width=width, height=height, scenario=EpsteinScenario(rng=seed, **params))
    agents = sorted(m.agents, key=lambda a: a.unique_id)
    sc = m.scenario
    st = ri.epstein_initial_state(random.Random(ri.seed_streams(seed)[1]), width, height, sc.citizen_density, sc.cop_density)
    assert np.array_equal(st["kind"], np.array([isinstance(a, Cop) for a in agents], dtype=np.uint8))
    assert np.array_equal(st["cell"], np.array([a.cell.coordinate[0] * height + a.cell.coordinate[1] for a in agents]))
    assert np.array_equal(st["hardship"], np.array([getattr(a, "hardship", 0.0) for a in agents]))
    assert np.array_equal(st["risk_aversion"], np.array([getattr(a, "risk_aversion", 0.0) for a in agents]))
    o = EpsteinOracle(width, height, st["kind"], st["cell"], st["hardship"], st["risk_aversion"], moore=sc.grid_type == "Moore",
                      vision=sc.cop_vision, legitimacy=sc.legitimacy, max_jail_term=sc.max_jail_term,
                      active_threshold=sc.active_threshold, arrest_prob_constant=sc.arrest_prob_constant, philox_seed=philox_seed)
    cit = st["kind"] == 0
    m.random.replay = True
    with replay.activation_context(Citizen, "step"), replay.activation_context(Cop, "step"):
        for step in range(8):
            m.step()
            o.step()
            assert np.array_equal(o.cell, np.array([a.cell.coordinate[0] * height + a.cell.coordinate[1] for a in agents])), step
            assert np.array_equal(o.state[cit], np.array([a.state.value for a in agents if isinstance(a, Citizen)], dtype=np.int8)), step
            assert np.array_equal(o.jail[cit], np.array([a.jail_sentence for a in agents if isinstance(a, Citizen)])), step
            assert o.counts() == (m.ACTIVE, m.QUIET, m.ARRESTED), step
    m.random.replay = False


@pytest.mark.parametrize("seed,params", [(2718, dict(initial_population=150)), (99, dict(initial_population=400, vision_max=3, metabolism_max=4))])
def test_sugarscape_from_a_seed(replay, seed, params):
    """Sugarscape G1MT with trade: initial traders from the seed (reference_init.sugarscape_initial_state), trajectory and
    every trade from the sequential restatement (oracle/sugarscape.py) == the unmodified model under the replayed order / draws."""
    from mesa.examples.advanced.sugarscape_g1mt.agents import Trader
    from mesa.examples.advanced.sugarscape_g1mt.model import SugarscapeG1mt, SugarScapeScenario
    from oracle.sugarscape import SugarscapeOracle

    philox_seed = seed + 1
    with replay.patched_model_random(philox_seed):
        m = SugarscapeG1mt(SugarScapeScenario(rng=seed, enable_trade=True, **params))
    W, H = m.grid.dimensions
    sc = m.scenario
    gen, stdlib_seed = ri.seed_streams(seed)
    st = ri.sugarscape_initial_state(random.Random(stdlib_seed), gen, W * H, sc.initial_population, sc.endowment_min, sc.endowment_max,
                                     sc.metabolism_min, sc.metabolism_max, sc.vision_min, sc.vision_max)
    agents = list(m.agents)
    cells = lambda: np.array([a.cell.coordinate[0] * H + a.cell.coordinate[1] for a in m.agents], dtype=np.int64)   # noqa: E731
    assert np.array_equal(st["cell"], cells())
    for name in ("sugar", "spice", "metabolism_sugar", "metabolism_spice", "vision"):
        assert np.array_equal(st[name], np.array([getattr(a, name) for a in agents])), name
    o = SugarscapeOracle(W, H, m.sugar_distribution, m.spice_distribution, np.arange(1, len(agents) + 1), st["cell"], st["sugar"],
                         st["spice"], st["metabolism_sugar"], st["metabolism_spice"], st["vision"], philox_seed=philox_seed)
    m.random.replay = True
    with replay.activation_context(Trader, "step"), replay.activation_context(Trader, "trade_with_neighbors"):
        for step in range(6):
            m.step()
            o.step()
            prices, partners = o.trade_phase()
            assert np.array_equal(o.uid, np.array([a.unique_id for a in m.agents])) and np.array_equal(o.cell, cells()), step
            assert np.array_equal(o.sugar, np.array([a.sugar for a in m.agents], dtype=np.float64)), step
            assert np.array_equal(o.spice, np.array([a.spice for a in m.agents], dtype=np.float64)), step
            assert [list(a.trade_partners) for a in m.agents] == partners and [list(a.prices) for a in m.agents] == prices, step
    m.random.replay = False


def test_neighbourhood_order_properties_the_dependency_graph_kernels_rely_on(replay):
    """csrc/epstein.cu / csrc/sugarscape.cu take a cell's radius-r neighbourhood from ONE offset table.  That is only right
    because, on the live reference grids, (1) a torus neighbourhood is translation invariant, (2) on a CLAMPED von Neumann
    grid it is the interior order with the absent cells dropped, (3) the radius-r list is a prefix of the radius-(r+1) list.
    (On clamped Moore grids (2) does NOT hold -- the Sugarscape model is von Neumann, the Epstein model a torus.)"""
    import random as _r

    from mesa.discrete_space import OrthogonalMooreGrid, OrthogonalVonNeumannGrid

    import torch  # noqa: F401  (mesa_b200.examples import torch)
    from mesa_b200.examples.sugarscape_g1mt import von_neumann_bfs_offsets
    from oracle.epstein import neighborhood_offsets

    def offsets(grid, x, y, r, wrap=None):
        out = []
        for c in grid._cells[(x, y)].get_neighborhood(radius=r):
            dx, dy = c.coordinate[0] - x, c.coordinate[1] - y
            if wrap:
                dx, dy = (dx + wrap[0] // 2) % wrap[0] - wrap[0] // 2, (dy + wrap[1] // 2) % wrap[1] - wrap[1] // 2
            out.append((dx, dy))
        return out

    for moore, klass in ((False, OrthogonalVonNeumannGrid), (True, OrthogonalMooreGrid)):
        torus = klass((23, 19), torus=True, random=_r.Random(1))
        for r in (1, 2, 4, 7):
            want = neighborhood_offsets(moore, r)
            for x, y in ((0, 0), (11, 9), (22, 18), (3, 17)):
                assert offsets(torus, x, y, r, wrap=(23, 19)) == want, (moore, r, x, y)          # (1)
            if r > 1:
                assert want[:len(neighborhood_offsets(moore, r - 1))] == neighborhood_offsets(moore, r - 1)   # (3)
    clamped = OrthogonalVonNeumannGrid((17, 21), torus=False, random=_r.Random(1))
    for r in (1, 3, 5, 7):
        table = [tuple(int(v) for v in o) for o in von_neumann_bfs_offsets(r)]
        assert table == neighborhood_offsets(False, r)
        for x in (0, 1, 4, 8, 15, 16):
            for y in (0, 2, 10, 19, 20):
                want = [(dx, dy) for dx, dy in table if 0 <= x + dx < 17 and 0 <= y + dy < 21]
                assert offsets(clamped, x, y, r) == want, (r, x, y)                              # (2)


def test_welfare_power_table_is_the_host_pow():
    """sugarscape_g1mt/agents.py:57-70 evaluates `sugar ** (ms / m_total)` on numpy float64 scalars: the table entries are the
    same bits, and the exact tie that CUDA's pow breaks (28 x 54 = 27 x 56 with equal metabolisms) holds in it."""
    from mesa_b200.examples.sugarscape_g1mt import welfare_power_table

    t = welfare_power_table(5, 200)
    for ms, mp, x in ((1, 1, 28), (5, 5, 54), (2, 3, 77), (5, 1, 199), (3, 4, 0), (4, 2, 1)):
        assert t[ms - 1, mp - 1, x] == np.float64(x) ** (np.int64(ms) / np.int64(ms + mp))
    assert t[4, 4, 28] * t[4, 4, 54] == t[4, 4, 27] * t[4, 4, 56]
