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


@pytest.mark.parametrize("seed,width,height,density,homophily,radius", [(2027, 15, 12, 0.75, 0.5, 1), (31337, 9, 21, 0.9, 0.7, 2)])
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
